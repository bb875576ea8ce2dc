#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Map a FASTQ (pair) through the reference's own per-read functions in-process
(oracle/_ref/libdartref.so, oracle/ref_taps.cpp) and print the SAM records; used where a test needs the
reference on a second index (the reference holds one index per process in globals).

usage: ref_replay.py <index prefix> <r1.fq> [<r2.fq>] [--mis N] [--max_dup N] [-m] [--all_sj]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402


def fastq(path):
    recs = []
    with open(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip(b"\n"); f.readline(); q = f.readline().rstrip(b"\n")
            name = h[1:].split(b" ")[0].split(b"/")[0].split(b"\t")[0].rstrip(b"\n")
            recs.append((name, s, q))
    return recs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("idx"); ap.add_argument("r1"); ap.add_argument("r2", nargs="?")
    ap.add_argument("--mis", type=int, default=0); ap.add_argument("--max_dup", type=int, default=100)
    ap.add_argument("-m", action="store_true"); ap.add_argument("--all_sj", action="store_true")
    a = ap.parse_args()
    R = po.Reference(a.idx)
    R.set_params(max_mismatch=a.mis, max_dup=a.max_dup, multi_hit=int(a.m), all_sj=int(a.all_sj), pair_end=int(bool(a.r2)))
    out = sys.stdout.buffer
    r1 = fastq(a.r1)
    if a.r2:
        for (n1, s1, q1), (n2, s2, q2) in zip(r1, fastq(a.r2)):
            out.write(R.map_pair(n1, s1, q1, n2, s2, q2))
    else:
        for n1, s1, q1 in r1:
            out.write(R.map_single(n1, s1, q1))


if __name__ == "__main__":
    main()
