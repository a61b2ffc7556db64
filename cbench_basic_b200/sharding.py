"""Image / tile partitioning over the GPUs of one node.

Images (and row-band tiles of a large latent) are independent coding units: every rank codes its own block
with no data-path collective.  The only exchange is the final gather of the per-unit stream byte counts
(8 bytes per unit) needed to lay the streams out in one container -- one all_gather over NCCL on GPUs (gloo in
the CPU tests).  lanes=1 reference-compatible streams code the whole batch in ONE stream and are therefore
single-GPU only.
"""
import struct
from typing import List, Sequence

import torch
import torch.distributed as dist


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int, local_rank: int = 0, local_world_size: int = 1) -> dict:
    """One process per GPU: pin the calling (main) thread -- and with it every thread and pinned allocation made later --
    to the CPUs of the NUMA node its GPU hangs off (/sys/bus/pci/devices/<bus id>/local_cpulist).  The stream of every
    coding call crosses PCIe through pinned host memory and is copied by a few host threads; with processes floating
    over both sockets the staging buffers land on the far node for half of the ranks.  Ranks that share a node split
    its CPUs between them.  Returns what was done (for logs); a no-op where sysfs has no topology (-1 / one node)."""
    import os
    info = {"bound": False}
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        info["pci"] = bus
        info["numa_node"] = int(open(base + "/numa_node").read())
        local = _parse_cpulist(open(base + "/local_cpulist").read())
        allowed = os.sched_getaffinity(0)
        cpus = sorted(local & allowed)
        info["allowed"] = len(allowed)
        info["local"] = len(cpus)
        if not cpus or len(cpus) == len(allowed):
            return info
        # ranks whose GPUs share the node take disjoint slices of it
        peers = []
        for d in range(torch.cuda.device_count()):
            q = torch.cuda.get_device_properties(d)
            qb = f"/sys/bus/pci/devices/{q.pci_domain_id:04x}:{q.pci_bus_id:02x}:{q.pci_device_id:02x}.0/numa_node"
            if int(open(qb).read()) == info["numa_node"]:
                peers.append(d)
        peers = peers[:max(1, local_world_size)] if device_index in peers[:max(1, local_world_size)] else peers
        if device_index in peers and len(cpus) >= 2 * len(peers):
            per = len(cpus) // len(peers)
            at = peers.index(device_index) * per
            cpus = cpus[at:at + per]
        os.sched_setaffinity(0, cpus)
        info["bound"] = True
        info["cpus"] = len(cpus)
    except Exception as e:  # no sysfs / no permission: run unbound
        info["error"] = repr(e)
    return info


def partition(n_units: int, world_size: int, rank: int) -> range:
    """Contiguous block of units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_units, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def row_band_tiles(H: int, n_tiles: int) -> List[range]:
    """Split H latent rows into n_tiles contiguous bands (cfg 5: 135 rows -> 8/16/32 bands).  Bands are coded as
    independent images: the context model pads band borders with zeros exactly like image borders."""
    n_tiles = max(1, min(n_tiles, H))
    return [partition(H, n_tiles, t) for t in range(n_tiles)]


def gather_sizes(local_sizes: Sequence[int], device=None, group=None) -> List[List[int]]:
    """all_gather of the per-unit stream lengths; returns one list per rank.  The single collective of the path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [list(local_sizes)]
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    count = torch.tensor([len(local_sizes)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    width = max(int(c.item()) for c in counts)
    mine = torch.zeros(max(width, 1), dtype=torch.int64, device=device)
    if len(local_sizes):
        mine[:len(local_sizes)] = torch.tensor(list(local_sizes), dtype=torch.int64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    return [allv[r][:int(counts[r].item())].tolist() for r in range(world)]


def unit_offsets(sizes_per_rank: Sequence[Sequence[int]]):
    """Byte offset of every unit inside the container body, in global unit order (rank-major)."""
    flat = [s for rank in sizes_per_rank for s in rank]
    offs, at = [], 0
    for s in flat:
        offs.append(at)
        at += s
    return flat, offs, at


def assemble_container(streams: Sequence[bytes]) -> bytes:
    """[u32 count][u32 len_i ...][stream_i ...], little endian."""
    head = struct.pack("<I", len(streams)) + b"".join(struct.pack("<I", len(s)) for s in streams)
    return head + b"".join(streams)


def split_container(data: bytes) -> List[bytes]:
    (count,) = struct.unpack_from("<I", data, 0)
    lens = struct.unpack_from(f"<{count}I", data, 4)
    at, out = 4 + 4 * count, []
    for n in lens:
        out.append(data[at:at + n])
        at += n
    if at != len(data):
        raise ValueError("container length does not match its directory")
    return out


def encode_sharded(encode_unit, n_units: int, rank: int = 0, world_size: int = 1, device=None, group=None):
    """Codes this rank's units with `encode_unit(unit_index) -> bytes`, gathers the sizes, and returns
    (local_streams, sizes_per_rank, this rank's byte offset inside the container body)."""
    mine = partition(n_units, world_size, rank)
    streams = [encode_unit(u) for u in mine]
    sizes = gather_sizes([len(s) for s in streams], device=device, group=group)
    _, offs, _ = unit_offsets(sizes)
    my_off = offs[mine.start] if len(mine) and mine.start < len(offs) else 0
    return streams, sizes, my_off
