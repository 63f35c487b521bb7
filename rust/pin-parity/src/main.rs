//! Prints one JSON document with the reference's own scores for `case<k>_ref.ppm` / `case<k>_dist.ppm` (k = 0, 1, ...)
//! found in the directory given as the first argument.  Every number goes through the same public functions
//! `EvalSession::calculate_metrics` calls (src/eval/session.rs:437-497), in the same order of conversions.
use codec_eval::metrics::{self, butteraugli, dssim, ssimulacra2, xyb};
use codec_eval::viewing::ViewingCondition;
use std::path::Path;

/// Binary PPM (P6, maxval 255) -> (width, height, tight RGB8)
fn read_ppm(path: &Path) -> Option<(usize, usize, Vec<u8>)> {
    let bytes = std::fs::read(path).ok()?;
    let mut pos = 0usize;
    let mut fields: Vec<String> = Vec::new();
    while fields.len() < 4 {
        while pos < bytes.len() && bytes[pos].is_ascii_whitespace() { pos += 1; }
        if pos < bytes.len() && bytes[pos] == b'#' {
            while pos < bytes.len() && bytes[pos] != b'\n' { pos += 1; }
            continue;
        }
        let start = pos;
        while pos < bytes.len() && !bytes[pos].is_ascii_whitespace() { pos += 1; }
        fields.push(String::from_utf8_lossy(&bytes[start..pos]).into_owned());
    }
    pos += 1; // the single whitespace byte after maxval
    assert_eq!(fields[0], "P6", "{}: not a binary PPM", path.display());
    assert_eq!(fields[3], "255", "{}: maxval must be 255", path.display());
    let (w, h): (usize, usize) = (fields[1].parse().ok()?, fields[2].parse().ok()?);
    let data = bytes[pos..].to_vec();
    assert_eq!(data.len(), w * h * 3, "{}: truncated", path.display());
    Some((w, h, data))
}

fn fnv1a64(data: &[u8]) -> u64 {
    data.iter().fold(0xcbf2_9ce4_8422_2325u64, |h, b| (h ^ u64::from(*b)).wrapping_mul(0x0000_0100_0000_01b3))
}

fn main() {
    let dir = std::env::args().nth(1).expect("usage: pin-parity <dir with case<k>_{ref,dist}.ppm>");
    let dir = Path::new(&dir);
    let mut rows = Vec::new();
    for k in 0.. {
        let (Some((w, h, r)), Some((w2, h2, t))) =
            (read_ppm(&dir.join(format!("case{k}_ref.ppm"))), read_ppm(&dir.join(format!("case{k}_dist.ppm"))))
        else { break };
        assert_eq!((w, h), (w2, h2));
        let psnr = metrics::calculate_psnr(&r, &t, w, h);
        let s2 = ssimulacra2::calculate_ssimulacra2(&r, &t, w, h).expect("ssimulacra2");
        let ds = dssim::calculate_dssim(&dssim::rgb8_to_dssim_image(&r, w, h), &dssim::rgb8_to_dssim_image(&t, w, h),
                                        &ViewingCondition::desktop()).expect("dssim");
        let ba = butteraugli::calculate_butteraugli(&r, &t, w, h).expect("butteraugli");
        let ba250 = butteraugli::calculate_butteraugli_with_intensity(&r, &t, w, h, 250.0).expect("butteraugli");
        let rt = xyb::xyb_roundtrip(&r, w, h);
        // with the XYB round trip applied to the reference first (MetricConfig::perceptual_xyb, session.rs:447-456)
        let s2_xyb = ssimulacra2::calculate_ssimulacra2(&rt, &t, w, h).expect("ssimulacra2");
        rows.push(format!(
            "  {{\"case\": {k}, \"width\": {w}, \"height\": {h}, \"psnr\": {psnr:e}, \"ssimulacra2\": {s2:e}, \
             \"dssim\": {ds:e}, \"butteraugli\": {ba:e}, \"butteraugli_250\": {ba250:e}, \
             \"ssimulacra2_xyb_ref\": {s2_xyb:e}, \"xyb_roundtrip_fnv1a64\": \"{:016x}\"}}", fnv1a64(&rt)));
    }
    // {:e} on f64 prints the shortest digits that round-trip, so the file carries the exact doubles
    println!("{{\"source\": \"codec-eval (fast-ssim2 / dssim-core / butteraugli as locked by its Cargo.lock)\",\n \"cases\": [\n{}\n ]}}",
             rows.join(",\n"));
}
