"""Data-parallel plumbing of the hot path.

Images are partitioned across ranks in contiguous blocks; nothing inside the
path communicates.  The only collective is the tail all-gather of the
fixed-capacity detection buffers: ONE kernel packs a rank's proposals and
counts into one buffer, ONE ``all_gather_into_tensor`` (NCCL over NVLink on
GPUs; gloo in the CPU tests) moves it - no eager tensor arithmetic in between."""
import torch
import torch.distributed as dist

from . import _lib as L


def shard_range(num_images, rank, world_size):
    """Contiguous block of images owned by `rank`: [lo, hi).  Remainder images go to the lowest ranks."""
    base, rem = divmod(num_images, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(proposals, num_valid, first_image_id, out=None):
    """(b,max_num,5) [x1,y1,x2,y2,score] + (b) int32 -> packed (b,max_num+1,6) f32 (mxd_pack_detections):
    row 0 of every image = [count, img_id, 0,0,0,0]; rows 1.. = [img_id, x1,y1,x2,y2,score], img_id = -1 on padding rows."""
    L.require_cuda(proposals, num_valid)
    b, m, _ = proposals.shape
    if out is None:
        out = torch.empty((b, m + 1, 6), dtype=torch.float32, device=proposals.device)
    L.call("mxd_pack_detections", L.dl(proposals.contiguous()), L.dl(num_valid.contiguous()), int(first_image_id), L.dl(out),
           L.current_stream(proposals.device))
    return out


def all_gather_packed(packed, group=None):
    """All-gather of equally sized packed blocks -> (world*b, max_num+1, 6), rank-major.  One collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return packed
    world = dist.get_world_size(group)
    out = torch.empty((world * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    return out


def unpack_detections(gathered):
    """Views into a gathered buffer: (detections (B,max_num,6) [img_id,x1,y1,x2,y2,score], counts (B) f32, image ids (B) f32)."""
    return gathered[:, 1:, :], gathered[:, 0, 0], gathered[:, 0, 1]


def gather_detections(proposals, num_valid, first_image_id=0, group=None):
    """pack -> one all-gather -> views.  Requires the same number of images on every rank (pad the shard otherwise).
    Returns (detections (world*b,max_num,6), counts (world*b) f32 - exact integers)."""
    dets, counts, _ = unpack_detections(all_gather_packed(pack_detections(proposals, num_valid, first_image_id), group))
    return dets, counts
