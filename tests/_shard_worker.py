"""Worker of tests/test_stream_sharding.py: one rank of a world_size-N gloo job.  Streams are sharded over the
ranks exactly like bench.py shards them over GPUs (rank r owns streams r*B .. r*B+B-1, no data-path collective);
the only collectives are the barrier and the MAX all-reduce of the step time.  The backend is the CPU oracle
(this is host-logic coverage; the GPU path has the same orchestration with CudaBackend)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_dir, B, n_frames = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    policy = sys.argv[4] if len(sys.argv) > 4 else "interval"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sys.path.insert(0, ROOT)
    from oracle.stream_backend import OracleBackend
    from revo_b200 import synth
    from revo_b200.stream import StreamTracker

    w, h = 160, 120
    cam = synth.intrinsics(w, h)
    seeds = [300 + rank * B + s for s in range(B)]
    fast = policy == "vote"                 # fast odd streams: their vote asks for keyframes
    streams = [synth.make_stream(sd, n_frames, w, h, max_trans=0.01 + 0.04 * (fast and sd % 2), max_rot_deg=0.5 + 2.5 * (fast and sd % 2))
               for sd in seeds]
    be = OracleBackend(cam, 3)
    st = StreamTracker(be, B, kf_interval=3, kf_policy=policy)
    frame = lambda i: (np.stack([s["frames"][i][0] for s in streams]), np.stack([s["frames"][i][1] for s in streams]))
    st.start(*frame(0))
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(1, n_frames):
        st.step(*frame(i))
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)           # timing: max over ranks
    frames = torch.tensor([float(B * (n_frames - 1))], dtype=torch.float64)
    dist.all_reduce(frames)                             # whole-job units
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), T_w_c=st.T_w_c, seeds=np.array(seeds), dt=dt.numpy(), frames=frames.numpy(),
             evals=st.total_evals, retracks=st.n_retracks)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
