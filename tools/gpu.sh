#!/usr/bin/env bash
# Rebuild librdm_b200.so and the C oracle HERE, then hand the command to gpurun (the built .so files travel with the snapshot; a stale
# library on the box cost a GPU run once).   tools/gpu.sh [--timeout S] [--gpus N] -- 'command'
set -e
cd "$(dirname "$0")/.."
python -c 'import __graft_entry__ as g; g.build()' > /dev/null
exec /usr/local/graft/bin/gpurun "$@"
