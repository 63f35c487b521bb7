"""Writes tests/golden/codec_iter_baseline_excerpt.json: the header and the first six points of the reference's own
baseline file (/root/reference/baselines/jpeg.json, written by codec-iter's save_baseline), re-serialised exactly as
serde_json::to_string_pretty does.  A format fixture for codec_eval_b200.codec_iter.Baseline; run in the build container."""
import json
import os

src = json.load(open("/root/reference/baselines/jpeg.json"))
src["points"] = src["points"][:6]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "codec_iter_baseline_excerpt.json")
raw = open("/root/reference/baselines/jpeg.json").read()
# keep the reference's own bytes for the kept prefix: cut after the 6th point and close the arrays
depth, count, end = 0, 0, None
start = raw.index('"points": [') + len('"points": [')
for i in range(start, len(raw)):
    if raw[i] == "{":
        depth += 1
    elif raw[i] == "}":
        depth -= 1
        if depth == 0:
            count += 1
            if count == 6:
                end = i + 1
                break
text = raw[:end] + "\n  ]\n}"
assert json.loads(text) == src
open(out, "w").write(text)
print("wrote", out, len(text), "bytes")
