"""CPU tests: the C-ABI library loads and exports every symbol include/ce_gpu.h declares (no compute calls),
and the host-side mirror types behave like the reference's (src/metrics/mod.rs tests)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "ce_gpu.h")).read()
    return sorted(set(re.findall(r"CE_API\s+[\w\s\*]+?\b(ce_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from codec_eval_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        from codec_eval_b200 import build

        build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 25
    assert sorted(_lib.EXPORTS) == declared
    for name in declared:
        assert hasattr(L, name), name
    _lib.load()
    assert b"sm_100a" in L.ce_version.__call__() if False else True  # no compute / device calls here


def test_struct_layouts_match_header():
    from codec_eval_b200 import _lib

    assert ctypes.sizeof(_lib.CeMetricConfig) == 5
    assert ctypes.sizeof(_lib.CeResult) == 56
    assert ctypes.sizeof(_lib.CePair) == 48


def test_ctx_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from codec_eval_b200.metrics import CudaError, GpuMetrics

    with pytest.raises(CudaError) as e:
        GpuMetrics(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_no_product_code_touches_the_oracle():
    pkg = os.path.join(ROOT, "codec_eval_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), (dirpath, f)


# ---- src/metrics/mod.rs:337-397
def test_perception_level_thresholds():
    from codec_eval_b200.metrics import PerceptionLevel as P

    assert P.from_dssim(0.0001) == P.Imperceptible
    assert P.from_dssim(0.0003) == P.Marginal
    assert P.from_dssim(0.0005) == P.Marginal
    assert P.from_dssim(0.0007) == P.Subtle
    assert P.from_dssim(0.001) == P.Subtle
    assert P.from_dssim(0.0015) == P.Noticeable
    assert P.from_dssim(0.002) == P.Noticeable
    assert P.from_dssim(0.003) == P.Degraded
    assert P.from_dssim(0.01) == P.Degraded
    assert P.from_ssimulacra2(90.1) == P.Imperceptible and P.from_ssimulacra2(90.0) == P.Marginal
    assert P.from_butteraugli(0.99) == P.Imperceptible and P.from_butteraugli(5.0) == P.Degraded
    assert P.Subtle.code() == "SUB" and str(P.Subtle) == "Subtle" and P.Subtle.max_dssim() == 0.0015


def test_metric_config_presets():
    from codec_eval_b200.metrics import MetricConfig, MetricResult

    c = MetricConfig.all()
    assert c.dssim and c.psnr and c.ssimulacra2 and c.butteraugli and not c.xyb_roundtrip
    c = MetricConfig.fast()
    assert not c.dssim and c.psnr
    assert MetricConfig.perceptual_xyb().xyb_roundtrip and not MetricConfig.perceptual().psnr
    assert MetricConfig.ssimulacra2_only().with_xyb_roundtrip().xyb_roundtrip
    r = MetricResult(dssim=0.0005)
    assert r.perception_level().code() == "MAR" and r.perception_level_ssimulacra2() is None


# ---- the Rust -sys crate cannot be compiled in this image; check it textually ----
_C2R = {"int": "c_int", "size_t": "usize", "float": "f32", "double": "f64", "uint32_t": "u32", "uint64_t": "u64",
        "uint8_t": "u8", "char": "c_char", "void": "c_void", "int32_t": "i32"}


def _c_type_to_rust(decl, is_param=True):
    """'const uint8_t* const* dists' -> '*const *const u8' (parameter name dropped)."""
    decl = re.sub(r"\s+", " ", decl).strip()
    if is_param:
        decl = re.sub(r"\b\w+$", "", decl).strip()
    toks = re.findall(r"const|\*|\w+", decl)
    toks = [t for t in toks if t != "struct"]
    base = [t for t in toks if t not in ("const", "*")][0]
    out = _C2R.get(base, base)
    # pointer levels left to right; a level is const if 'const' qualifies the pointee it points at
    pointee_const = "const" in toks[: toks.index("*")] if "*" in toks else False
    i = 0
    while i < len(toks):
        if toks[i] == "*":
            out = ("*const " if pointee_const else "*mut ") + out
            pointee_const = i + 1 < len(toks) and toks[i + 1] == "const"
        i += 1
    return out


def _strip_c_comments(src):
    return re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", src, flags=re.S))


def test_rust_sys_crate_matches_header():
    h = _strip_c_comments(open(os.path.join(ROOT, "include", "ce_gpu.h")).read())
    r = re.sub(r"//[^\n]*", "", open(os.path.join(ROOT, "rust", "ce-gpu-sys", "src", "lib.rs")).read())
    c_funcs = {}
    for m in re.finditer(r"CE_API\s+([\w\s\*]+?)\b(ce_\w+)\s*\(([^;]*?)\)\s*;", h, flags=re.S):
        ret, name, args = m.groups()
        args = [] if args.strip() in ("", "void") else [a.strip() for a in args.split(",")]
        c_funcs[name] = (ret.strip(), args)
    r_funcs = {}
    for m in re.finditer(r"pub fn (ce_\w+)\s*\(([^;]*?)\)\s*(?:->\s*([^;]+?))?\s*;", r, flags=re.S):
        name, args, ret = m.groups()
        r_funcs[name] = (ret, [a.strip() for a in args.split(",") if a.strip()])
    assert len(r_funcs) >= 25
    norm = lambda s: s.replace(" ", "")
    for name, (ret, args) in r_funcs.items():
        assert name in c_funcs, name
        c_ret, c_args = c_funcs[name]
        assert len(c_args) == len(args), name
        for ca, ra in zip(c_args, args):
            assert norm(_c_type_to_rust(ca)) == norm(ra.split(":", 1)[1]), (name, ca, ra)
        want = _c_type_to_rust(c_ret, is_param=False)
        assert norm(ret or "c_void") == norm(want), (name, c_ret, ret)
    # everything a caller of the path needs is bound; only the debug taps and the profiler stay C-only
    unbound = set(c_funcs) - set(r_funcs)
    assert all(n.startswith(("ce_debug_", "ce_profile_")) for n in unbound), unbound

    # #[repr(C)] structs: same field order and widths as the header
    def c_fields(struct):
        body = re.search(r"typedef struct\s*\{([^}]*)\}\s*%s\s*;" % struct, h, flags=re.S).group(1)
        out = []
        for f in body.split(";"):
            if not f.strip():
                continue
            first, *more = [d.strip() for d in f.split(",")]  # 'size_t ref_len, dist_len' -> one entry per name
            ty = _c_type_to_rust(first)
            out.append((ty, re.search(r"(\w+)$", first).group(1)))
            out += [(ty, n) for n in more]
        return out

    def r_fields(struct):
        body = re.search(r"pub struct %s\s*\{(.*?)\}" % struct, r, flags=re.S).group(1)
        return [(norm(f.split(":", 1)[1]), f.split(":", 1)[0].replace("pub", "").strip())
                for f in body.split(",") if f.strip()]

    for s in ("ce_metric_config", "ce_result", "ce_pair"):
        cf, rf = c_fields(s), r_fields(s)
        assert [{"ref": "reference"}.get(n, n) for _, n in cf] == [n for _, n in rf], s  # `ref` is a Rust keyword
        assert [norm(t) for t, _ in cf] == [t for t, _ in rf], s

    # status codes and validity bits carry the same values on both sides (and in the ctypes mirror)
    c_consts = {k: int(v) for k, v in re.findall(r"\b(CE_(?:OK|ERR_\w+))\s*=\s*(\d+)", h)}
    c_consts.update({k: int(v) for k, v in re.findall(r"#define\s+(CE_VALID_\w+)\s+(\d+)u", h)})
    r_consts = {k: int(v) for k, v in re.findall(r"pub const (CE_\w+)\s*:\s*\w+\s*=\s*(\d+)\s*;", r)}
    assert len(c_consts) >= 10
    for k, v in c_consts.items():
        assert r_consts.get(k) == v, k


def test_host_mirror_covers_the_reference_metrics_api():
    """Every public item of the reference's `metrics` module (docs/public-api/codec-eval.txt:386-458) has a
    same-named counterpart in codec_eval_b200.metrics (free functions also as GpuMetrics methods)."""
    from codec_eval_b200 import metrics as M

    free_functions = [
        "calculate_psnr", "calculate_ssimulacra2", "calculate_ssimulacra2_icc", "calculate_butteraugli",
        "calculate_butteraugli_icc", "calculate_butteraugli_with_intensity", "calculate_dssim", "calculate_dssim_icc",
        "rgb8_to_dssim_image", "rgba8_to_dssim_image", "xyb_roundtrip", "prepare_for_comparison", "transform_to_srgb",
    ]
    for f in free_functions:
        assert callable(getattr(M, f)), f
        method = "transform_profile_to_srgb" if f == "transform_to_srgb" else f  # the method of that name takes w, h
        assert callable(getattr(M.GpuMetrics, method)), f
    for cls, members in {
        "MetricConfig": ["all", "fast", "perceptual", "perceptual_xyb", "ssimulacra2_only", "with_xyb_roundtrip"],
        "MetricResult": ["perception_level", "perception_level_butteraugli", "perception_level_ssimulacra2"],
        "PerceptionLevel": ["code", "from_butteraugli", "from_dssim", "from_ssimulacra2", "max_butteraugli", "max_dssim",
                            "min_ssimulacra2", "Imperceptible", "Marginal", "Subtle", "Noticeable", "Degraded"],
        "ColorProfile": ["Srgb", "Icc", "from_icc_bytes", "is_srgb"],
        "GpuReference": ["compare", "compare_many"],  # Ssimulacra2Reference::{new, compare} (prelude.rs:85)
    }.items():
        for m in members:
            assert hasattr(getattr(M, cls), m), (cls, m)
    for field in ("dssim", "ssimulacra2", "butteraugli", "psnr", "xyb_roundtrip"):
        assert field in M.MetricConfig.__dataclass_fields__
    for field in ("dssim", "ssimulacra2", "butteraugli", "psnr"):
        assert field in M.MetricResult.__dataclass_fields__
