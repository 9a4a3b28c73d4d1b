#!/bin/bash
# Installs the UNMODIFIED reference quant path into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it
# travels to the GPU box, where /root/reference does not exist).  bench.py --impl reference and the
# cpu_baseline leg import it from there (tests/golden/ref_loader.py, QASR_REFERENCE_ROOT) and time the
# reference's own fake-quant forward on the box's host cores.
#
# The reference is a pure-Python NeMo fork whose package does not import here (pytorch_lightning, hydra,
# omegaconf, onnx ... are absent), so `pip install --target baseline/_ref /root/reference`
# would install a package that cannot be imported.  What the quant path needs is six files, byte for byte:
#   nemo/__init__.py, nemo/package_info.py          (package shell; NEMO_PACKAGE_BUILDING=1 skips the rest)
#   nemo/quantization/utils/{quant_utils,quant_modules,quantize_model}.py
#   nemo/collections/asr/parts/{jasper,activations}.py
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -d "$SRC/nemo/quantization/utils" ]; then
  echo "install_ref: reference tree not found at $SRC" >&2
  exit 1
fi
rm -rf "$DST"
for f in nemo/__init__.py nemo/package_info.py \
         nemo/quantization/utils/quant_utils.py nemo/quantization/utils/quant_modules.py \
         nemo/quantization/utils/quantize_model.py \
         nemo/collections/asr/parts/jasper.py nemo/collections/asr/parts/activations.py; do
  mkdir -p "$DST/$(dirname "$f")"
  cp "$SRC/$f" "$DST/$f"
done
( cd "$SRC" && sha256sum nemo/__init__.py nemo/package_info.py nemo/quantization/utils/*.py \
    nemo/collections/asr/parts/jasper.py nemo/collections/asr/parts/activations.py ) > "$DST/SHA256SUMS.src"
( cd "$DST" && sha256sum -c --quiet <(grep -v distill_data SHA256SUMS.src) )
echo "installed reference quant path into $DST"
