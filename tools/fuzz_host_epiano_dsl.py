"""TEST TOOL: the built-in fuzzer's random ElectricPiano scripts (tools/fuzz_host_builtin.py) played on the kernel GENERATED from
tests/graphs/electric_piano.graph, executed on the host, against the oracle bank.  OSCEN_B200_JIT_ARRAY_SMEM picks the home of the
array nodes' state.   PYTHONPATH=tests:. python tools/fuzz_host_epiano_dsl.py FIRST LAST"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ('tests', '', 'tools'):
    sys.path.insert(0, os.path.join(ROOT, d))
import numpy as np
import oracle_lib as O
from host_kernel import HostEngine
import fuzz_host_builtin as F
text = open(os.path.join(ROOT, 'tests', 'graphs', 'electric_piano.graph')).read()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
ok = bad = 0
for seed in range(lo, hi):
    n, vpi, script, x = F.make_script(O.K_EPIANO, seed)
    (vr, ir), (vh, ih) = F.play(O.OracleBank(O.K_EPIANO, n, vpi), script, x), F.play(HostEngine(text, n), script, x)
    if np.array_equal(vr, vh) and np.array_equal(ir, ih): ok += 1
    else:
        bad += 1; print("MISMATCH seed", seed, float(np.nanmax(np.abs(vr - vh))))
print("electric_piano.graph generated kernel (array mode %s) vs oracle bank: ok %d, mismatch %d" % (os.environ.get("OSCEN_B200_JIT_ARRAY_SMEM", "default 2"), ok, bad))
