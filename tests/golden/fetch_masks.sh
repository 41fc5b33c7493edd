#!/bin/sh
# Copies the land-mask DATA files (not source) used by the parity tests from the reference's data/mkmask/
# into tests/golden/masks/ so that the GPU box (which has no /root/reference) can run the same cases.
# Format: src/ocean/topo.F90:41-64.
set -e
REF=${REF:-/root/reference}
for f in mask_natl8 test6x6x4 test6x12x4_2 mask_gateway mask_global_96x38x12 mask_global_48x19x4 mask_natl16; do
  cp "$REF/data/mkmask/$f" "$(dirname "$0")/masks/$f"
done
