"""Runs the reference's OWN test scripts, unmodified, in the build container (needs /root/reference):
  (a) on the stand-ins of oracle/clshim (the reference's kernels compiled for the host), and
  (b) on the product's PyOpenCL-signature binding (synchrad_b200/compat) with the CPU emulation of the CUDA
      kernels in place of the library call (there is no GPU here; on a B200 the same binding calls the library).
The scripts assert nothing; they print the deviation of the integrated energy from the analytic undulator estimate
(tests/test_undulator_analytic.py:78-89).  usage: python tools/run_reference_tests.py [far|near|both] > profiles/...
"""
import os
import re
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from oracle import run_reference as rr   # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'far'
scripts = {'far': 'tests/test_undulator_analytic.py', 'near': 'tests/test_undulator_analytic_near.py'}
from emu import emu   # noqa: E402
emu.build()
for key in (('far', 'near') if which == 'both' else (which,)):
    path = os.path.join(rr.REFERENCE_ROOT, scripts[key])
    for backend in ('clshim', 'compat_emu'):
        t0 = time.time()
        out = rr.run_script(path, backend=backend, seed=0)
        dev = re.findall(r'Deviation from analytic estimate is ([0-9.]+)%', out)
        modes = re.findall(r'Running (.*)', out)
        print(f'== {scripts[key]} on {backend} (numpy seed 0, {time.time() - t0:.0f} s)')
        for ln in out.splitlines():
            if ln.startswith(('Running', 'Platform', 'Compiler', 'Done', 'Deviation', 'WARNING', '  ')):
                print('   ', ln)
        assert len(dev) == 2, out[-2000:]
