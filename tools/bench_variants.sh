#!/bin/bash
# development aid: IVF C2 scan-kernel time under different knobs (NDB_SCAN_MODE / QT / NW / SEG)
run() {
  echo -n "$*: "
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f QPS step %.3f ms kernel %.3f ms e2e %.0f QPS recall %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['recall_at_10']))"
}
for qt in 4 8; do for nw in 8 4; do
  run NDB_SCAN_MODE=direct NDB_SCAN_QT=$qt NDB_SCAN_NW=$nw NDB_IVF_SEG_BLOCKS=64
  echo "C1 QT=$qt NW=$nw"; NDB_SCAN_MODE=direct NDB_SCAN_QT=$qt NDB_SCAN_NW=$nw python tools/quick_bench.py c1 | tail -3
done; done
