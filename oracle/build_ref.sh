#!/bin/bash
# Stage the UNMODIFIED reference (cassiePython/NeRF-Art) for the GPU box: /root/reference does not exist there, oracle/_ref/ travels
# with the gpurun snapshot (git-ignored, not gpurun-ignored).  The reference is pure Python, so "building" it is copying its sources
# where they lie -- *.py, the YAML configs, criteria/neg_text.txt and one camera file (70 KB; images are NOT copied, tests synthesise
# them) -- into oracle/_ref/reference/.  Nothing under oracle/_ref/ is tracked or imported by the product; users:
#   bench.py --impl reference   (the reference's own volume_render, timed on the host cores and, as `reference_gpu`, on the B200)
#   tests/test_dropin_gpu.py    (the reference's render.py / train.py driven unchanged on top of the nerfart_b200 mirror)
# usage: oracle/build_ref.sh [reference root, default /root/reference]
set -euo pipefail
SRC=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
DST=$HERE/_ref/reference
[ -f "$SRC/train.py" ] || { echo "no reference tree at $SRC (fine on the GPU box: oracle/_ref ships prebuilt)"; exit 0; }
rm -rf "$DST"; mkdir -p "$DST"
(cd "$SRC" && find . \( -name '*.py' -o -name '*.yaml' -o -name 'neg_text.txt' \) -not -path './.git/*' -print0 | tar --null -T - -cf -) | tar -xf - -C "$DST"
mkdir -p "$DST/data/fangzhou_nature"
cp "$SRC/data/fangzhou_nature/cameras.npz" "$DST/data/fangzhou_nature/cameras.npz"
chmod -R u+w "$DST"
(cd "$SRC" && find . \( -name '*.py' -o -name '*.yaml' -o -name 'neg_text.txt' \) -not -path './.git/*' | sort | xargs sha256sum) > "$HERE/_ref/SHA256SUMS"
echo "staged $(find "$DST" -type f | wc -l) files ($(du -sh "$DST" | cut -f1)) in $DST"
