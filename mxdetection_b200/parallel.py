"""Data-parallel plumbing of the hot path.

Images are partitioned across ranks in contiguous blocks; nothing inside the
path communicates.  The only collective is the tail all-gather of the
fixed-capacity detection buffers (NCCL on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(num_images, rank, world_size):
    """Contiguous block of images owned by `rank`: [lo, hi).  Remainder images go to the lowest ranks."""
    base, rem = divmod(num_images, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(proposals, num_valid, first_image_id):
    """(b,max_num,5) [x1,y1,x2,y2,score] + (b) -> (b,max_num,6) [img_id,x1,y1,x2,y2,score]; padding rows img_id=-1."""
    b, m, _ = proposals.shape
    ids = torch.arange(first_image_id, first_image_id + b, device=proposals.device, dtype=torch.float32)
    ids = ids[:, None].expand(b, m).clone()
    rows = torch.arange(m, device=proposals.device)[None, :]
    ids[rows >= num_valid[:, None].to(rows.dtype)] = -1.0
    return torch.cat([ids[..., None], proposals], dim=2)


def gather_detections(proposals, num_valid, first_image_id=0, group=None):
    """All-gather of equally sized per-rank detection blocks -> ((world*b,max_num,6), (world*b) int32).

    Requires the same number of images on every rank (pad the shard otherwise)."""
    packed = pack_detections(proposals, num_valid, first_image_id).contiguous()
    counts = num_valid.to(torch.int32).contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return packed, counts
    world = dist.get_world_size(group)
    out = torch.empty((world * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    cnt = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out, packed, group=group)   # rank-major concatenation along dim 0
    dist.all_gather_into_tensor(cnt, counts, group=group)
    return out, cnt
