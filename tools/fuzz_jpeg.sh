#!/usr/bin/env bash
# Memory-safety fuzz of the JPEG decoding and affine-warp code under AddressSanitizer + UBSan (CPU, test infrastructure).
#   bash tools/fuzz_jpeg.sh [SECONDS] [JOBS]
set -eu
here=$(cd "$(dirname "$0")/.." && pwd)
secs=${1:-60}; jobs=${2:-4}
work=$(mktemp -d)
python - "$work" <<'PY'
import sys, numpy as np, cv2
rng = np.random.default_rng(0)
out = sys.argv[1]
k = 0
for (H, W) in ((64, 64), (37, 53), (8, 8), (128, 96)):
    for img in (rng.integers(0, 256, size=(H, W), dtype=np.uint8), cv2.GaussianBlur(rng.integers(0, 256, size=(H, W), dtype=np.uint8), (0, 0), 2.0)):
        for params in ([cv2.IMWRITE_JPEG_QUALITY, 95], [cv2.IMWRITE_JPEG_QUALITY, 30, cv2.IMWRITE_JPEG_OPTIMIZE, 1],
                       [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 4]):
            open("%s/seed%02d.jpg" % (out, k), "wb").write(cv2.imencode(".jpg", img, params)[1].tobytes())
            k += 1
PY
g++ -O1 -g -fwrapv -fsanitize=address,undefined -fno-sanitize-recover=all -I "$here/recursion_cellular_image_classification_b200/csrc" \
    "$here/tests/jpeg_fuzz.cpp" "$here/tests/jpeg_host.cpp" -o "$work/jpeg_fuzz"
g++ -O1 -g -fwrapv -ffp-contract=off -fsanitize=address,undefined -fno-sanitize-recover=all -fno-sanitize=float-cast-overflow \
    -I "$here/recursion_cellular_image_classification_b200/csrc" "$here/tests/warp_fuzz.cpp" "$here/tests/warp_host.cpp" -o "$work/warp_fuzz"
"$work/warp_fuzz" 1 "$(( secs / 4 + 1 ))"
pids=()
for j in $(seq 1 "$jobs"); do "$work/jpeg_fuzz" "$j" "$secs" "$work"/seed*.jpg & pids+=($!); done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
exit $rc
