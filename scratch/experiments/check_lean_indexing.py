#!/usr/bin/env python
"""Index logic of scratch/experiments/track_lean.cu restated in Python: every point of a level is visited exactly once, the
cached / uncached segments partition a thread's points, the shared-memory slots of different threads never collide, and
the two-slot software pipeline of `segment` retires its points in order.  (No GPU needed.)"""
import itertools


def thread_points(n, C, T, pcap, member, tid):
    stride = C * T
    first_idx = member * T + tid
    my_iter = (n - first_idx + stride - 1) // stride if first_idx < n else 0
    my_cached = min(my_iter, pcap)
    cached = [first_idx + k * stride for k in range(my_cached)]
    tail = [first_idx + k * stride for k in range(my_cached, my_iter)]
    return cached, tail


def segment_order(k0, k1):
    """Retirement order of the A/B pipeline in `segment` (indices k0..k1-1)."""
    if k0 >= k1:
        return []
    out, nxt = [], k0
    A, nxt = nxt, nxt + 1
    left = k1 - k0 - 1
    B = None
    while True:
        if left > 0:
            B, nxt = nxt, nxt + 1
        out.append(A)
        if left <= 0:
            break
        if left > 1:
            A, nxt = nxt, nxt + 1
        out.append(B)
        if left <= 1:
            break
        left -= 2
    return out


def main():
    for n, C, T, pcap in itertools.product((0, 1, 127, 128, 1000, 1024, 1025, 20281, 30000), (1, 2, 8), (128, 256), (0, 3, 18)):
        seen = []
        for member in range(C):
            for tid in range(T):
                cached, tail = thread_points(n, C, T, pcap, member, tid)
                assert len(cached) <= pcap
                seen += cached + tail
                # shared-memory slot of cached point k of thread tid: byte offset (k*3 + c)*T*4 + 4*tid, c = 0..2
                slots = {(k * 3 + c) * T * 4 + 4 * tid for k in range(len(cached)) for c in range(3)}
                assert len(slots) == 3 * len(cached) and (not slots or max(slots) < pcap * T * 12)
        assert sorted(seen) == list(range(n)), (n, C, T, pcap)
    for m in range(0, 12):
        assert segment_order(5, 5 + m) == list(range(5, 5 + m)), m
    print("ok")


if __name__ == "__main__":
    main()
