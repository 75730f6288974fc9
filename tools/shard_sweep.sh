#!/bin/bash
# 2^24 sharded proof over N real ranks (NCCL) for several block / tail settings: tools/shard_sweep.sh N "lb:tail lb:tail ..."
N=$1; shift
for cfg in $@; do
  lb=${cfg%%:*}; tl=${cfg##*:}
  rm -f /tmp/zk_id_$cfg
  for r in $(seq 0 $((N-1))); do
    ZK_SHARD_LOG_BLOCK=$lb ZK_SHARD_TAIL_LOG=$tl timeout 300 python tools/shard_nccl_worker.py --rank $r --world $N --log-n 24 --id-file /tmp/zk_id_$cfg --out gpurun_out/shard_${N}_${lb}_${tl}_$r.json --reps 5 > /dev/null 2>&1 &
  done
  wait
  python - <<PY
import json
d = json.load(open("gpurun_out/shard_${N}_${lb}_${tl}_0.json"))
print("N=$N lb=$lb tail=$tl", d["transcript_sha256"][:8], {k: round(v, 3) for k, v in d["phase_ms"].items()}, "first_tail", d["first_tail"], flush=True)
PY
done
