"""The device cardinality search, modelled in Python (tools/select_model.py mirrors the digit arithmetic of
k_yl_spec / k_sel_begin / k_radix_hist / radix_pick_block), against a plain sort: threshold key, tie quota and tie count
must come out right for every guess — the speculation may only change the number of passes."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import select_model as sm  # noqa: E402


@pytest.mark.parametrize("KB,dt,it", [(32, np.float32, np.uint32), (64, np.float64, np.uint64)])
def test_select_model_matches_sort(KB, dt, it):
    rng = np.random.default_rng(KB)
    for trial in range(20):
        n = int(rng.integers(5, 400))
        v = (rng.standard_normal(n) * 10 ** rng.uniform(-3, 3)).astype(dt)
        if trial % 3 == 0:
            v[rng.integers(0, n, n // 2)] = v[0]          # ties
        if trial % 4 == 0:
            v[rng.integers(0, n, n // 3)] = 0
        keys = np.abs(v).view(it)
        k = int(rng.integers(1, n))
        srt = np.sort(keys)[::-1]
        thr = int(srt[k - 1])
        for guess in (0, thr, thr ^ 1, thr + 5000, int(srt[min(k + 3, n - 1)]), int(srt[0])):
            for dbits, spec in ((11, True), (11, False), (8, False)):
                st, passes = sm.select(keys, k, KB, guess, dbits, spec)
                assert st["prefix"] == thr
                assert st["count_eq"] == int(np.sum(keys == thr))
                assert st["k_rem"] == k - int(np.sum(keys > thr))
            if KB == 32:
                assert sm.select(keys, k, KB, thr)[1] == 0            # a right guess decides all three levels
                assert sm.select(keys, k, KB, thr ^ 1)[1] <= 1        # right to 22 bits: one pass left
