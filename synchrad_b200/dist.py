"""The one exchange step of the path: sum of the per-rank spectra onto rank 0.

Replaces `SynchRad._gather_result_mpi` (calc.py:560-571: barrier + MPI Reduce(SUM, MPI.DOUBLE) per
radiation key + reduce(total_weight)) with ONE `torch.distributed.reduce` of the stacked float64
spectra (NCCL over NVLink on GPUs; gloo in the CPU tests) and one for the scalars.  As in the
reference, non-root ranks end with zero arrays and `total_weight = None`.
"""
import torch


def reduce_to_root(comm, tensors, total_weight, counters=None):
    """tensors: list of float64 tensors of one shape (device of the process group's backend).
    Returns (tensors, total_weight, counters) as seen by this rank after the reduction."""
    rank = comm.get_rank()
    buf = torch.stack(tensors)
    dev = buf.device
    if comm.get_backend() == 'gloo' and buf.is_cuda:      # ranks sharing one GPU (tests): gloo reduces host tensors
        buf = buf.cpu()
    comm.reduce(buf, dst=0, op=comm.ReduceOp.SUM)
    n_cnt = 0 if counters is None else counters.numel()
    scal = torch.zeros(1 + n_cnt, dtype=torch.float64, device=buf.device)
    scal[0] = total_weight
    if counters is not None:
        scal[1:] = counters.to(torch.float64)       # exact below 2^53
    comm.reduce(scal, dst=0, op=comm.ReduceOp.SUM)
    if rank == 0:
        out = list(buf.to(dev).unbind(0))
        tw = float(scal[0].item())
        cnt = scal[1:].to(torch.int64) if counters is not None else None
    else:
        out = [torch.zeros_like(t) for t in tensors]
        tw = None
        cnt = torch.zeros_like(counters) if counters is not None else None
    return out, tw, cnt
