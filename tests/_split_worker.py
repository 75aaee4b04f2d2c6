"""Worker of tests/test_gpu_split.py: one rank (= one GPU) of the single-pair edge split (BASELINE config 5).
Every rank builds the same pyramids, the mailbox handles are all-gathered, then the ranks track in lock-step with
one 32-double exchange per evaluation over NVLink (inside the persistent kernel)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_dir, w, h = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    from revo_b200 import api, synth

    p = synth.make_pair(5, w, h, xi=synth.XI_CONFIG1)
    fx, fy, cx, cy, _, _ = p["cam"]
    st = api.ImgPyramidSettings(PYR_MIN_LVL=2, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)
    ctx = api.Context(local)
    key = api.ImgPyramidRGBD(ctx, st, None, *p["key"])
    cur = api.ImgPyramidRGBD(ctx, st, None, *p["cur"])
    key.makeKeyframe()
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    # single-GPU result of this rank (all ranks must agree with it up to summation order)
    s1, R1, T1, e1 = trk.trackFrames(np.eye(3), np.zeros(3), key, cur)
    ev1 = list(trk.last_result.n_evals)
    blob = trk.splitExport(rank, world)
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    trk.splitOpen(b"".join(blobs))
    dist.barrier()
    absent = os.environ.get("REVO_SPLIT_TEST_ABSENT_RANK")
    if absent is not None:
        # watchdog test: one rank never launches; the others must return REVO_ERR_COMM after the watchdog period
        comm_rc = 0
        if rank != int(absent):
            try:
                trk.trackFramesSplit(np.eye(3), np.zeros(3), key, cur)
            except api.RevoError as e:
                comm_rc = e.code
        np.savez(os.path.join(out_dir, f"split{rank}.npz"), comm_rc=np.array(comm_rc))
        dist.barrier()
        dist.destroy_process_group()
        return
    s2, R2, T2, e2 = trk.trackFramesSplit(np.eye(3), np.zeros(3), key, cur)
    ev2 = list(trk.last_result.n_evals)
    # a second launch re-uses the mailboxes (sequence numbers must keep them apart)
    s3, R3, T3, e3 = trk.trackFramesSplit(np.eye(3), np.zeros(3), key, cur)
    np.savez(os.path.join(out_dir, f"split{rank}.npz"), R1=R1, T1=T1, R2=R2, T2=T2, R3=R3, T3=T3, e=np.array([e1, e2, e3]),
             ev1=np.array(ev1), ev2=np.array(ev2), s=np.array([s1, s2, s3]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
