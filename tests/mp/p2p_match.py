"""2-rank check (run under torchrun on >= 2 GPUs): a LightGlue match whose OLD keyframe lives in the other rank's
feature store (pulled over NVLink through the CUDA-IPC mapping) must equal the match against a locally extracted copy
of the same keyframe, and every rank's bank must hold all ranks' descriptors in global frame order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                      # noqa: E402
from d_vins_b200 import capi, sharding            # noqa: E402
from oracle import synth                          # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == 2
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    wpath = bench.make_weights() if rank == 0 else None
    dist.barrier()
    wpath = bench.make_weights()
    H, W, b = 480, 752, 2
    eng = capi.Engine(device=local, height=H, width=W, max_batch=b, max_vio=160, weights_path=wpath, bank_capacity=256,
                      store_capacity=64, world_size=world, rank=rank)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0])
    st = synth.Stream(H, W, period=12, margin=64)

    def extract(ids, frame_of):
        frames = np.stack([st.frame(frame_of(int(t))) for t in ids])
        vio = np.zeros((b, 160, 2), np.float32)
        for i, t in enumerate(ids):
            vio[i, :150] = synth.vio_points(150, H, W, 500 + frame_of(int(t)))
        eng.batch_upload(frames)
        eng.batch_extract(vio, np.full((b,), 150, np.int32), ids)
        first = eng.batch_commit(b)
        return first

    globals_seen = []
    for R in range(2):
        ids = sharding.round_frame_ids(R, rank, world, b)
        first = extract(ids, lambda t: t)
        assert first == int(ids[0]), (first, ids)
        globals_seen.extend(eng.batch_read_global(i) for i in range(b))
    # bank rows are in global frame order on every rank
    bank = eng.bank_export()
    assert bank.shape[0] == 2 * world * b
    mine = [int(t) for R in range(2) for t in sharding.round_frame_ids(R, rank, world, b)]
    for g, t in zip(globals_seen, mine):
        assert np.array_equal(bank[t], g)
    gathered = [None, None]
    dist.all_gather_object(gathered, bank.tobytes())
    assert gathered[0] == gathered[1], "banks differ between ranks"
    # remote old keyframe: rank 0 matches its frame 4 against frame 2 (rank 1's); rank 1 its frame 6 against frame 0
    q = int(sharding.round_frame_ids(1, rank, world, b)[0])
    old_remote = int(sharding.round_frame_ids(0, 1 - rank, world, b)[0])
    assert sharding.owner_of(old_remote, world, b) == 1 - rank
    m_remote, s_remote = eng.batch_match(np.array([q], np.int64), np.array([old_remote], np.int64))[0]
    # local copy of the same keyframe under ids 40.. (every rank takes part in the round's all-gather)
    copy_ids = np.array([40 + rank * b, 41 + rank * b], np.int64)
    extract(copy_ids, lambda t: old_remote if t == copy_ids[0] else old_remote + 1)
    m_local, s_local = eng.batch_match(np.array([q], np.int64), copy_ids[:1])[0]
    assert np.array_equal(m_remote, m_local) and np.array_equal(s_remote, s_local), (len(m_remote), len(m_local))
    print("rank %d: remote match == local match (%d pairs), bank replicated (%d rows)" % (rank, len(m_remote), bank.shape[0]),
          flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
