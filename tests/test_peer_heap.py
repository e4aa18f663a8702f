"""The allocator of the peer windows (multi_stark_b200/csrc/peer_heap.hpp) on the CPU: the row-sharded prover relies on every rank
computing the SAME (segment, offset) for a block from the same sequence of calls, so the allocator must be a pure function of
that sequence. A small driver is compiled with g++ and replayed against a Python model of "first fit at the lowest offset with
coalescing"; two independent instances must agree call for call."""
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
#include "peer_heap.hpp"
#include <cstdio>
#include <vector>
int main() {
    size_t first, bytes;
    if (scanf("%zu %zu", &first, &bytes) != 2) return 2;
    msg::FirstFitHeap a(first, bytes), b(first, bytes);
    char op;
    size_t v;
    while (scanf(" %c %zu", &op, &v) == 2) {
        if (op == 'a') {
            size_t x = a.alloc(v), y = b.alloc(v);
            if (x != y) { printf("DIVERGED\n"); return 1; }
            if (x == msg::FirstFitHeap::npos) printf("a -1\n"); else printf("a %zu\n", x);
        } else {
            bool x = a.free(v), y = b.free(v);
            if (x != y) { printf("DIVERGED\n"); return 1; }
            printf("f %d\n", x ? 1 : 0);
        }
    }
    printf("s %zu %zu %zu\n", a.live_blocks(), a.free_blocks(), a.free_bytes());
    return 0;
}
'''


class Model:
    ALIGN = 512

    def __init__(self, first, size):
        self.free = {first: size - first} if size > first else {}
        self.used = {}

    def alloc(self, n):
        need = max((n + self.ALIGN - 1) // self.ALIGN * self.ALIGN, self.ALIGN)
        for off in sorted(self.free):
            sz = self.free[off]
            if sz >= need:
                del self.free[off]
                if sz > need:
                    self.free[off + need] = sz - need
                self.used[off] = need
                return off
        return -1

    def release(self, off):
        if off not in self.used:
            return 0
        sz = self.used.pop(off)
        self.free[off] = sz
        merged = {}
        for o in sorted(self.free):  # coalesce adjacent free blocks
            if merged and (lo := max(merged)) + merged[lo] == o:
                merged[lo] += self.free[o]
            else:
                merged[o] = self.free[o]
        self.free = merged
        return 1


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    d = tmp_path_factory.mktemp("peer_heap")
    src, exe = d / "driver.cpp", d / "driver"
    src.write_text(DRIVER)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "multi_stark_b200", "csrc"), str(src), "-o", str(exe)])
    return str(exe)


@pytest.mark.parametrize("seed,first,size", [(1, 4096, 1 << 20), (2, 0, 1 << 16), (3, 4096, 4096 + 512), (4, 0, 1 << 24)])
def test_first_fit_heap_matches_model_and_is_deterministic(driver, seed, first, size):
    rng = random.Random(seed)
    m = Model(first, size)
    script, want, live = ["%d %d" % (first, size)], [], []
    for _ in range(3000):
        if live and rng.random() < 0.45:
            off = live.pop(rng.randrange(len(live)))
            script.append("f %d" % off)
            want.append("f %d" % m.release(off))
        elif rng.random() < 0.03:
            bogus = rng.randrange(size) | 1  # never a block offset (offsets are multiples of 512)
            script.append("f %d" % bogus)
            want.append("f %d" % m.release(bogus))
        else:
            n = rng.choice([1, 8, 511, 512, 513, 4096, 100000, rng.randrange(1, max(2, size // 8))])
            off = m.alloc(n)
            script.append("a %d" % n)
            want.append("a %d" % off)
            if off >= 0:
                assert off % 512 == 0 and off >= first and off + n <= size
                live.append(off)
    want.append("s %d %d %d" % (len(m.used), len(m.free), sum(m.free.values())))
    out = subprocess.run([driver], input="\n".join(script) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    assert [x for x in out if x] == want
    # everything released: one free block again (full coalescing)
    for off in list(m.used):
        m.release(off)
    assert len(m.free) <= 1 and sum(m.free.values()) == max(size - first, 0)
