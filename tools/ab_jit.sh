for n in 65536 1048576; do
for mb in 8 10; do for h in 0 1; do
  echo "voices=$n minblocks=$mb hoist=$h"
  OSCEN_BENCH_WAYS=gen OSCEN_B200_JIT_MIN_BLOCKS=$mb OSCEN_B200_JIT_ADSR_HOIST=$h python tools/bench_dsl.py 20 $n fm_voice 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('   %.4f ms  %.3g v-s/s' % (d['kernel_ms'], d['voice_samples_per_s']))"
done; done; done
