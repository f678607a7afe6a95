# how do the slot pitch / scratch offset inside the arena change the transform passes?  (log_n_rows = 20, stage times)
for cfg in "0 0" "0 64" "0 1024" "0 4096" "64 0" "1024 0" "4096 64" "32 32" "2048 2048" "520 264"; do
  set -- $cfg
  S2C_TILE_PAD_KB=$1 S2C_SCRATCH_OFF_KB=$2 python profiles/stage_times.py 20 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pad_kb=$1 off_kb=$2', d['sha'], d['total'], {k:d[k] for k in ('ifft_low','fft_mid','fft_low','trace_merkle_leaves','constraints')})"
done
