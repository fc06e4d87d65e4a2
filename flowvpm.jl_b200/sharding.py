"""One-process-per-GPU sharding of the P2P sweeps.

Targets are block-sharded across ranks; every rank needs all sources, so each
sweep is preceded by ONE all-gather of the 8 x N source buffer ([x y z rho Gx Gy
Gz sigma] per particle, the reference's fmm source-buffer layout,
src/FLOWVPM_fmm.jl:62-71) and, for the SFS sweep, one all-gather of the 9 x N
velocity gradients.  torch.distributed (NCCL over NVLink on GPUs, gloo in the
CPU tests) is the plumbing; the pair arithmetic is libvpm_cuda's device entry
points (vpm_uj_device / vpm_sfs_device).
"""
import torch
import torch.distributed as dist


def shard_size(n, world):
    return (n + world - 1) // world


def shard_bounds(n, world, rank):
    """half-open target range [t0, t1) owned by `rank`"""
    c = shard_size(n, world)
    t0 = min(n, rank * c)
    return t0, min(n, t0 + c)


def pad_local(local, n, world, rank, pad_row):
    """local shard (rows = particles) padded to the common shard size with `pad_row`"""
    c = shard_size(n, world)
    t0, t1 = shard_bounds(n, world, rank)
    assert local.shape[0] == t1 - t0
    if local.shape[0] == c:
        return local
    pad = pad_row.to(local).expand(c - local.shape[0], local.shape[1])
    return torch.cat([local, pad], dim=0)


def all_gather_rows(local_padded, world):
    """[c, k] per rank -> [world*c, k], rank-major == particle order"""
    if world == 1:
        return local_padded
    full = torch.empty((world * local_padded.shape[0], local_padded.shape[1]), dtype=local_padded.dtype,
                       device=local_padded.device)
    dist.all_gather_into_tensor(full, local_padded.contiguous())
    return full


# a padding particle: zero strength (contributes exactly 0 to every sum), sigma = 1
PAD_SRC8 = torch.tensor([[0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]], dtype=torch.float64)
PAD_J9 = torch.zeros((1, 9), dtype=torch.float64)


class ShardedField:
    """Device-resident shard of a particle field: rows [t0, t1) of the 8 x N source
    buffer, their U/J (12 per particle) and SFS (3 per particle)."""

    def __init__(self, handle, src8_local, n_total, rank, world, kernel_id):
        self.h = handle
        self.n = int(n_total)
        self.rank, self.world = rank, world
        self.kernel_id = int(kernel_id)
        self.t0, self.t1 = shard_bounds(self.n, world, rank)
        self.c = shard_size(self.n, world)
        self.src8_local = pad_local(src8_local, self.n, world, rank, PAD_SRC8).contiguous()
        dev = self.src8_local.device
        self.out12 = torch.zeros((self.c, 12), dtype=torch.float64, device=dev)
        self.out3 = torch.zeros((self.c, 3), dtype=torch.float64, device=dev)
        self.full8 = None
        self.launches = 0

    def gather_sources(self):
        self.full8 = all_gather_rows(self.src8_local, self.world)
        return self.full8

    def uj(self, flags=0):
        """one U/J sweep for this rank's targets (all-gather + kernel), stream-ordered"""
        full = self.gather_sources()
        stream = torch.cuda.current_stream().cuda_stream
        nt_total = self.world * self.c  # padded particles included: they are harmless
        self.h.check(self.h.lib.vpm_uj_device(self.h.ptr, full.data_ptr(), nt_total, self.rank * self.c,
                                              (self.rank + 1) * self.c, self.out12.data_ptr(),
                                              self.kernel_id, flags, stream))
        self.launches += 3
        return self.out12

    def sfs(self, flags=0):
        """SFS sweep: needs the final J of every particle -> second all-gather"""
        J_local = self.out12[:, 3:12].contiguous()
        fullJ = all_gather_rows(J_local, self.world)
        stream = torch.cuda.current_stream().cuda_stream
        nt_total = self.world * self.c
        self.h.check(self.h.lib.vpm_sfs_device(self.h.ptr, self.full8.data_ptr(), fullJ.data_ptr(), None,
                                               nt_total, self.rank * self.c, (self.rank + 1) * self.c,
                                               self.out3.data_ptr(), self.kernel_id, flags, stream))
        self.launches += 3
        return self.out3
