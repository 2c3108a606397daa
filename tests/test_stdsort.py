"""CPU test: basal_b200/csrc/stdsort.cuh (libstdc++'s std::sort restated over an index array, used on the device where
SortHits4PE's std::sort is not stable — SURVEY trap 10) against the real std::sort, on arrays full of equal keys."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define HD inline
#define __host__
#define __device__
#define __forceinline__
#include "stdsort.cuh"
struct E { unsigned key; unsigned id; };
static int check(std::vector<E> &v) {
    const int n = (int)v.size();
    std::vector<E> w = v;
    std::sort(w.data(), w.data() + n, [](const E &a, const E &b) { return a.key < b.key; });
    std::vector<uint16_t> p(n); for (int i = 0; i < n; i++) p[i] = (uint16_t)i;
    stdsort(p.data(), n, [&](uint16_t a, uint16_t b) { return v[a].key < v[b].key; });
    for (int i = 0; i < n; i++) if (v[p[i]].id != w[i].id) return 1;
    return 0;
}
int main() {
    srand(7); int bad = 0;
    for (int trial = 0; trial < 60000; trial++) {
        int n = 1 + rand() % (trial % 50 == 0 ? 2000 : 120);
        int kinds = 1 + rand() % (1 + rand() % (2 * n));
        std::vector<E> v(n); for (int i = 0; i < n; i++) v[i] = {(unsigned)(rand() % kinds), (unsigned)i};
        if (trial % 7 == 0) std::sort(v.begin(), v.end(), [](const E &a, const E &b) { return a.key < b.key; });
        if (trial % 11 == 0) std::reverse(v.begin(), v.end());
        if (trial % 13 == 0) for (int i = 0; i < n; i++) v[i].key = (i * 7919u) % (unsigned)kinds;
        bad += check(v);
    }
    for (int n : {17, 64, 100, 257, 1000, 2048}) {
        std::vector<E> v(n); for (int i = 0; i < n; i++) v[i] = {(unsigned)(i < n / 2 ? i : n - i) / 3, (unsigned)i};
        bad += check(v);
    }
    printf("bad=%d\n", bad); return bad != 0;
}
'''


def test_stdsort_matches_libstdcxx(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = str(tmp_path / "t")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "basal_b200", "csrc"), str(src), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout + out.stderr
