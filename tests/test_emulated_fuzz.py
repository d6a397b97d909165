"""Differential fuzz of the kernel logic (CPU emulation of srb_core.cuh) against the oracle."""
import contextlib
import io

import numpy as np
import pytest

import fuzzcases
from emu import emu


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_random_problems_match_oracle(oracle, seed):
    rs = np.random.RandomState(seed)
    for i in range(20):
        A, tracks, dt, kw = fuzzcases.rand_case(rs)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = oracle.calculate_spectrum(A, tracks, dt, **kw)
        kinds = ['direct'] if A.get('Features') or A['grid'][-1][0] < 2 else ['direct', 'recur', 'drec']
        if 'recur' in kinds and A.get('mode', 'far') == 'far':
            kinds.append('pair')
            if not kw['comp'].startswith('spheric'):
                kinds.append('pair_ws')     # warp-specialised form (transverse basis, 4- and 8-node tiles)
        for kind in kinds:
            for nPC in (1, 3):
                with contextlib.redirect_stdout(io.StringIO()):
                    rad, cnt = emu.run(A, tracks, dt, kind=kind, nPC=nPC, **kw)
                e = fuzzcases.vector_errors(rad, ref['radiation'])
                assert e < 1e-9, (seed, i, kind, nPC, e, A['grid'], A.get('mode'), A.get('Features'), kw)
