"""Clip-range sharding across GPUs (one process per GPU, no data-path collective).

Mirrors main_func's split (PredictAndGenerate.py:274-275,300-306): worker i owns the contiguous frame
range [start + i*step, min(end, start + (i+1)*step)), step = ceil(n / workers), and restarts the
clip state (depth history, range EMA) at its `begin` — so an N-GPU run is bit-identical to the
reference run with Num_Workers = N, not to a single-worker run.  Sub-clip names keep the reference's
`{begin}_{end}` convention (PredictAndGenerate.py:243) so Check_Clips / Combine_Clips still apply.

torch.distributed is used only for the rendezvous, the barrier around timed regions and the
max-over-ranks of timings (NCCL on GPUs, gloo in the CPU tests).
"""
import os

from . import tables


def world_from_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0,1,2,3,8,10,11} (the sysfs cpulist format)."""
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def bind_near_gpu(pci_bus_id, min_cpus=4, sysfs="/sys/bus/pci/devices"):
    """Restrict this process (and the threads it starts later: the library's copy pool) to the CPUs of the NUMA node
    the GPU hangs off, so that pinned staging buffers are first-touched next to the GPU's PCIe root.  One process
    per GPU is the reference's own layout (main_func spawns a worker per clip range); the end-to-end path at 4-8
    GPUs is host-memory bound, which is where placement matters.  `pci_bus_id` as CUDA prints it
    ("0000:1b:00.0").  Returns the CPU set applied, or None when sysfs has no answer or fewer than `min_cpus` of the
    node's CPUs are available to this process (nothing is changed then)."""
    try:
        with open(os.path.join(sysfs, pci_bus_id.lower(), "local_cpulist")) as f:
            local = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if len(target) < min_cpus or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        return target
    except (OSError, ValueError, AttributeError):
        return None


def shard_for_rank(start_frame, end_frame, video_length, world, rank):
    """(begin, end) of this rank's clip range, or None when there are more ranks than ranges."""
    ranges = tables.clip_ranges(start_frame, end_frame, video_length, world)
    return ranges[rank] if rank < len(ranges) else None


def flush_ranges(begin, end, video_length, max_frame_count):
    """The sub-clips one worker writes, exactly as nibba_woka's loop names them (PredictAndGenerate.py:221-250):
    [(last_i, i, frames)] -> file f"{last_i}_{i}.mp4" holding `frames` frames.  Frame i-1 is appended at iteration i
    (one-frame look-ahead) and the final frame at the last iteration, so a full sub-clip "0_15" holds frames 0..14
    and the next one is named from 16.  A range of exactly one frame writes NOTHING: the reference's progress print
    divides by `step_taken = i - begin` (:237-238), which is zero when the first flush happens at i == begin; the
    ZeroDivisionError lands in the worker's catch-all handler (:259-272), which logs it and returns before ffmpeg is
    started (probed by running the unmodified loop, tests/test_reference_callers.py)."""
    stop = min(end, video_length)
    if stop - begin == 1:
        return []
    out, pending, last_i = [], 0, begin
    for i in range(begin, stop):
        if i != begin:
            pending += 1
        if i == stop - 1:
            pending += 1
        if pending == max_frame_count or i == min(end - 1, video_length - 1):
            out.append((last_i, i, pending))
            last_i, pending = i + 1, 0
    return out


def subclip_name(first, last):
    return f"{first}_{last}.mp4"


class Ranks:
    """Barrier + max-over-ranks; degenerates to no-ops for a single process."""

    def __init__(self, backend=None, device=None):
        self.rank, self.world, self.local_rank = world_from_env()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group(backend or "nccl")
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, value):
        if self.dist is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device or "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, obj):
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
