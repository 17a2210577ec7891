"""CPU: the stream-K schedule of the Gram kernel (host logic, csrc/schedule.h) covers every (tile, stage) exactly
once, keeps the segments of a tile contiguous (fixed reduction order) and balances weighted work across CTAs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bayesianlinearregressors.jl_b200", "csrc")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("sched") / "schedule_test"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", CSRC, os.path.join(ROOT, "tests", "cpp", "schedule_test.cpp"), "-o", str(out)],
                   check=True)
    return str(out)


@pytest.mark.parametrize("nt,n_stages,G,wd", [
    (8, 1 << 19, 148, 38),   # cfg3: D=1024, N=2^24 in 32-observation stages, hybrid tiling's diagonal weight
    (8, 1 << 20, 148, 40),   # same in 16-observation stages with the 2x4 tiling's weight
    (2, 32768, 148, 38),     # cfg2: D=256, N=2^20
    (2, 65536, 148, 40),
    (32, 131072, 148, 38),   # cfg5: D=4096, N=2^22
    (32, 262144, 148, 36),
    (1, 1, 148, 40), (1, 7, 148, 40), (3, 5, 148, 64), (2, 1000, 1, 40), (5, 999, 7, 50), (8, 131072, 132, 44),
])
def test_schedule_properties(exe, nt, n_stages, G, wd):
    res = subprocess.run([exe, str(nt), str(n_stages), str(G), str(wd)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    tag, nseg, mx, mn = res.stdout.split()
    assert tag == "ok"
    T = nt * (nt + 1) // 2
    assert int(nseg) <= G + T
    total = n_stages * (nt * wd + (T - nt) * 64)
    # no CTA carries more than its fair share plus one stage of the heaviest tile per segment boundary
    assert int(mx) <= total // G + 2 * 64 + 1


@pytest.mark.parametrize("nt,n_stages,G,wd,flush", [(8, 128, 148, 40, 32), (2, 512, 148, 40, 32), (32, 32, 148, 40, 48), (3, 5, 148, 64, 32),
                                                    (8, 96, 132, 44, 64)])
def test_schedule_with_flush_cost(exe, nt, n_stages, G, wd, flush):
    """Periodic variant: multi-segment CTAs are charged a flush per segment; coverage / order invariants still hold
    and the heaviest CTA (work + flushes) stays within a few stages of the mean."""
    res = subprocess.run([exe, str(nt), str(n_stages), str(G), str(wd), str(flush)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    tag, nseg, mx, mn = res.stdout.split()
    assert tag == "ok"
    T = nt * (nt + 1) // 2
    total = n_stages * (nt * wd + (T - nt) * 64)
    assert int(mx) <= total // G + flush * (T // G + 3) + 3 * 64


@pytest.mark.parametrize("MI", list(range(1, 17)))
def test_mid_team_deal_is_a_balanced_partition(tmp_path, MI):
    """csrc/mid_deal.h: the two warps of a K1m team (gram_mid.cu; MI = 9 .. 12 block rows in production) partition the lower
    triangle and carry the same number of 8 x 8 sub-tiles up to one."""
    out = tmp_path / "mid_deal_test"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", CSRC, os.path.join(ROOT, "tests", "cpp", "mid_deal_test.cpp"), "-o", str(out)], check=True)
    res = subprocess.run([str(out), str(MI)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    tag, c0, c1 = res.stdout.split()
    assert tag == "ok" and int(c0) + int(c1) == MI * (MI + 1) // 2
    assert abs(int(c0) - int(c1)) <= 1, (MI, c0, c1)
