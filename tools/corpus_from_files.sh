#!/bin/bash
# Config 3 from wav FILES through the CLI a user runs: N ten-minute PCM_16 clips in /dev/shm, detected by
# `torchrun -m softspoken_b200.corpus` on G GPUs (files sharded per rank, reader thread per rank, one NCCL gather),
# and the same list on one GPU; the two CSVs must be identical.  Usage: tools/corpus_from_files.sh <n_gpus> [n_clips]
set -e
G=${1:-2}; CLIPS=${2:-32}
D=$(mktemp -d -p /dev/shm ss_corpus_XXXX)
trap 'rm -rf "$D"' EXIT
python - "$D" "$CLIPS" <<'PY'
import sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
from softspoken_b200 import synth, wavio
d, n = sys.argv[1], int(sys.argv[2])
base = [synth.synth_pcm16(600.0, s) for s in range(4)]
with open(os.path.join(d, "files.txt"), "w") as f:
    for i in range(n):
        p = os.path.join(d, f"clip{i:03d}.wav")
        wavio.write_wav_pcm16(p, np.roll(base[i % 4], 997 * i), 22050)
        f.write(p + "\n")
PY
python -m softspoken_b200.corpus "$D/files.txt" "$D/one.csv" --max-batch 1005 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$G" --master-addr 127.0.0.1 --master-port 29547 \
    -m softspoken_b200.corpus "$D/files.txt" "$D/multi.csv" --max-batch 1005 | tail -2
cmp "$D/one.csv" "$D/multi.csv" && echo "CSV of $(wc -l < "$D/one.csv") lines identical for 1 and $G ranks"
