"""N-rank check (run under torchrun on >= 2 GPUs): a LightGlue match whose OLD keyframe lives in another rank's
feature store (read over NVLink through the CUDA-IPC mapping, seqlock-verified) must equal the match against a locally
extracted copy of the same keyframe; every rank's bank must hold all ranks' descriptors in global frame order; a
keyframe nobody holds must be reported per pair.  The assertions themselves are bench.multi_gpu_preflight - the same
function bench.py runs before timing at --gpus >= 2 (its result is the bench line's `parity_multi`)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                      # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world >= 2
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    wpath = bench.make_weights() if rank == 0 else None
    dist.barrier()
    wpath = bench.make_weights()
    res = bench.multi_gpu_preflight(dist, rank, world, local, wpath)
    assert res["banks_equal"], res
    assert res["remote_eq_local"], res
    assert res["missing_keyframe_is_per_pair"], res
    print("rank %d: remote match == local match, bank replicated: %s" % (rank, json.dumps(res)), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
