"""CPU test of the partitioned build's host logic: the ownership map (recommendersystems_b200/csrc/ownmap.h) deals every
segment of the node range over the ranks in contiguous, balanced pieces that cover every node exactly once."""
import os
import subprocess

from conftest import ROOT

SRC = r'''
#include <cstdio>
#include <vector>
#include "ownmap.h"
int main() {
    const long long cases[][4] = {{1000, 9000, 200, 8}, {7, 3, 0, 8}, {1, 1, 1, 2}, {5000000, 45000000, 0, 8}, {12, 0, 0, 5}};
    for (auto& c : cases) {
        OwnMap m;
        m.parts = (int)c[3];
        m.n_segs = 0;
        m.seg[0] = 0;
        for (int k = 0; k < 3; k++)
            if (c[k]) { m.seg[m.n_segs + 1] = m.seg[m.n_segs] + c[k]; m.n_segs++; }
        const long long n = m.seg[m.n_segs];
        const long long step = n > 2000000 ? 997 : 1;            // sample the big case
        for (int s = 0; s < m.n_segs; s++) {
            std::vector<long long> cnt(m.parts, 0);
            int prev = 0;
            for (long long i = m.seg[s]; i < m.seg[s + 1]; i += step) {
                const int r = own_rank(m, i);
                if (r < 0 || r >= m.parts || r < prev) { printf("bad owner %d of node %lld\n", r, i); return 1; }   // contiguous, ascending
                prev = r;
                cnt[r]++;
            }
            if (step == 1) {
                long long lo = cnt[0], hi = cnt[0];
                for (long long x : cnt) { lo = x < lo ? x : lo; hi = x > hi ? x : hi; }
                if (hi - lo > 1) { printf("unbalanced segment %d: %lld..%lld\n", s, lo, hi); return 1; }
            }
        }
    }
    printf("ownmap ok\n");
    return 0;
}
'''


def test_ownership_map_partitions_every_segment(tmp_path):
    src = tmp_path / "own.cpp"
    src.write_text(SRC)
    exe = tmp_path / "own"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "recommendersystems_b200", "csrc"),
                           str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "ownmap ok" in out.stdout, out.stdout
