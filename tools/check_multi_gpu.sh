#!/bin/bash
# BASELINE config 3 evidence: the detections CSV of an N-rank run (files sharded per GPU, NCCL gather to rank 0) is
# byte-identical to the single-GPU CSV.  Usage: tools/check_multi_gpu.sh <n_gpus> [n_clips]
set -e
N=${1:-2}; CLIPS=${2:-9}
D=$(mktemp -d)
python - "$D" "$CLIPS" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from softspoken_b200 import synth, wavio
d, n = sys.argv[1], int(sys.argv[2])
with open(os.path.join(d, "files.txt"), "w") as f:
    for i in range(n):
        p = os.path.join(d, f"clip{i:02d}.wav")
        wavio.write_wav_pcm16(p, synth.synth_pcm16(15.0 + 11.0 * (i % 4), 100 + i), 22050)
        f.write(p + "\n")
PY
python -m softspoken_b200.corpus "$D/files.txt" "$D/one.csv" --max-batch 32
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29533 \
    -m softspoken_b200.corpus "$D/files.txt" "$D/multi.csv" --max-batch 32
cmp "$D/one.csv" "$D/multi.csv" && echo "config3 OK: $(wc -l < "$D/one.csv") CSV lines identical for 1 and $N ranks"
head -3 "$D/one.csv"
