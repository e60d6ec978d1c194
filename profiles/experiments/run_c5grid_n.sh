#!/bin/bash
# C5 grid mode at N GPUs with the given library builds: profiles/experiments/run_c5grid_n.sh N [lib ...]
N=$1; shift
for lib in "default" "$@"; do
  if [ "$lib" = default ]; then unset NDSB200_LIB; else export NDSB200_LIB=$lib; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --workload c5_grid --strides 100 > /tmp/c5g.out 2> /tmp/c5g.err
  python - "$lib" <<'P'
import json, sys
lines = [l for l in open("/tmp/c5g.out") if l.startswith("{")]
if not lines:
    print(sys.argv[1], "NO OUTPUT"); print(open("/tmp/c5g.err").read()[-1500:]); sys.exit(0)
d = json.loads(lines[-1]); c = d["config"]
print(sys.argv[1].split("/")[-1], "%.4g" % d["value"], "us/stride %.1f" % (1e3 * c["ms_per_stride"]),
      "xshare %.2f" % c["exchange_share"], "b2b %.4g" % c["back_to_back"]["value"], c["exchange_check"])
P
done
