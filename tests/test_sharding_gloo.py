"""World-size-2 gloo test of the N>1 host logic (block sharding of targets + all-gather of
the 8 x N source buffer / 9 x N gradients) on CPU.  The per-shard pair arithmetic is done
by the ORACLE here (test infrastructure) because there is no GPU; on the B200 box the same
sharding module drives libvpm_cuda (tests/test_multigpu_gpu.py, bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vpm_import import load
    from oracle import oracle
    vpm = load()
    from flowvpm_jl_b200 import sharding
    pf = vpm.fields.cloud_field(n, kernel=vpm.winckelmans, seed=33)
    src8 = vpm.source_system_to_buffer(pf)                    # 8 x n
    t0, t1 = sharding.shard_bounds(n, world, rank)
    local = torch.from_numpy(np.ascontiguousarray(src8[:, t0:t1].T))
    padded = sharding.pad_local(local, n, world, rank, sharding.PAD_SRC8)
    full = sharding.all_gather_rows(padded, world)            # [world*c, 8]
    c = sharding.shard_size(n, world)
    assert full.shape == (world * c, 8)
    full_np = np.asfortranarray(full.numpy().T)
    # every rank must see the same, correctly ordered source buffer
    assert np.array_equal(full_np[:, :n], src8)
    assert np.all(full_np[4:7, n:] == 0)
    # this rank's targets against all (padded) sources
    tb = np.zeros((16, t1 - t0), order="F")
    tb[0:3] = src8[0:3, t0:t1]
    oracle.direct_buffers(tb, 0, t1 - t0, full_np, 0, world * c, "winckelmans")
    # second exchange: gradients
    J_local = torch.from_numpy(np.ascontiguousarray(tb[7:16].T))
    J_full = sharding.all_gather_rows(sharding.pad_local(J_local, n, world, rank, sharding.PAD_J9), world)
    np.save(os.path.join(out_dir, f"uj_{rank}.npy"), tb[4:16])
    if rank == 0:
        np.save(os.path.join(out_dir, "J_full.npy"), J_full.numpy()[:n].T)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    n, world = 301, 2   # odd: the last shard is padded
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from vpm_import import load
    from oracle import oracle
    vpm = load()
    pf = vpm.fields.cloud_field(n, kernel=vpm.winckelmans, seed=33)
    ref = pf.particles.copy(order="F")
    oracle.uj_direct(ref, n, "winckelmans")
    got = np.concatenate([np.load(tmp_path / f"uj_{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(got[0:3], ref[9:12, :n])      # same per-target source order: bit-identical
    assert np.array_equal(got[3:12], ref[15:24, :n])
    assert np.array_equal(np.load(tmp_path / "J_full.npy"), ref[15:24, :n])
