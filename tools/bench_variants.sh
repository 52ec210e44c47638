#!/bin/bash
# development aid: IVF C2 scan-kernel time under different tile / segment knobs
for nw in 7 3; do for seg in 8 16 32 64; do
  echo -n "NW=$nw SEG=$seg: "
  NDB_SCAN_NW=$nw NDB_IVF_SEG_BLOCKS=$seg python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f QPS step %.3f ms kernel %.3f ms e2e %.0f QPS recall %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['recall_at_10']))"
done; done
