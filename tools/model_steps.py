#!/usr/bin/env python3
"""CPU model of the cell kernel's work on a uniform random frame (no GPU): tiles, stream steps, distance tests and the
hit rate for the plan mb_plan_describe returns (or an overridden one).  Used to compare tile shapes offline."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from test_plan_host import describe, TRIC

def model(pl, n=1_000_000, seed=1, pairs=364146133.5, cw=64):
    fd = pl['fd']; hx = pl['hx']; rows = pl['rows']
    rng = np.random.default_rng(seed)
    f = rng.random((n, 3))
    c = (f * fd).astype(np.int64)
    cell = c[:, 0] + fd[0] * (c[:, 1] + fd[1] * c[:, 2])
    cnt = np.bincount(cell, minlength=int(np.prod(fd))).reshape(fd[2], fd[1], fd[0])
    csx = np.concatenate([np.zeros((fd[2], fd[1], 1), np.int64), np.cumsum(cnt, axis=2)], axis=2)  # prefix along x
    tdx = fd[0] // hx
    # home tile populations
    nh = cnt.reshape(fd[2], fd[1], tdx, hx).sum(3)   # [z][y][tile]
    T = np.zeros_like(nh)
    fx = np.arange(tdx) * hx
    T += nh  # self run
    for dy, dz, lo, hi in rows:
        # candidate count of the row for every tile: periodic wrap along all dims
        xs = np.arange(lo, hi + 1)
        for x in xs:
            T += np.roll(np.roll(np.roll(cnt, -dz, 0), -dy, 1), -x, 2)[:, :, fx]
    nz = nh > 0
    batches = (nh + 31) // 32
    steps_per_batch = (T + cw - 1) // cw
    steps = (batches * steps_per_batch)[nz].sum()
    # tests: per batch nh4 homes x cw candidates per step
    nh4 = ((np.minimum(nh, 32) + 3) // 4) * 4  # approx for single-batch tiles
    tests = (nh4 * steps_per_batch * cw)[nz].sum()
    real = (nh * T)[nz].sum()
    return dict(tiles=int(nz.sum()), mean_nh=float(nh[nz].mean()), mean_T=float(T[nz].mean()), steps=int(steps),
                home_steps=int((nh4 * steps_per_batch)[nz].sum()), tests=float(tests), real_tests=float(real),
                hit=pairs / tests, hit_real=pairs / real)

if __name__ == '__main__':
    pl = describe(TRIC.astype(np.float64), 1.2, 1_000_000)
    print({k: v for k, v in pl.items() if k != 'rows'})
    print(model(pl))
