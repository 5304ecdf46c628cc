"""Multi-GPU plumbing for the alignment path: one process per GPU (torch.distributed), reads partitioned across
ranks by estimated DP cost, reference set broadcast once, variable-length result strings gathered to every
rank.  No collective sits inside the DP: every (read, reference range) job is independent."""
import numpy as np


def partition_by_cost(costs, world_size):
    """Contiguous partition of items into world_size shards with near-equal summed cost.
    Returns a list of (begin, end) index pairs, one per rank (possibly empty)."""
    n = len(costs)
    total = float(sum(costs))
    bounds = [0]
    acc = 0.0
    k = 1
    for i, c in enumerate(costs):
        acc += c
        while k < world_size and acc >= total * k / world_size and len(bounds) < world_size:
            bounds.append(i + 1)
            k += 1
    while len(bounds) < world_size:
        bounds.append(n)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def broadcast_references(refs, dist, device, src=0):
    """refs: list of (name, sequence) on the source rank (ignored elsewhere).  Broadcasts one packed uint8
    buffer plus an offset table through the process group (NCCL over NVLink on the GPU box, gloo in the CPU
    tests) and returns the list of (name, sequence) on every rank."""
    import torch
    rank = dist.get_rank()
    if rank == src:
        blob = '\n'.join('%s\t%s' % (n, s) for n, s in refs).encode()
        size = torch.tensor([len(blob)], dtype=torch.int64, device=device)
    else:
        blob = b''
        size = torch.zeros(1, dtype=torch.int64, device=device)
    dist.broadcast(size, src=src)
    n = int(size.item())
    if rank == src:
        buf = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).to(device)
    else:
        buf = torch.empty(n, dtype=torch.uint8, device=device)
    dist.broadcast(buf, src=src)
    text = bytes(buf.cpu().numpy().tobytes()).decode()
    out = []
    for line in text.split('\n'):
        if line:
            name, seq = line.split('\t', 1)
            out.append((name, seq))
    return out


def all_gather_strings(strings, dist, device):
    """Gathers each rank's list of result strings to every rank (lengths first, then one padded byte tensor),
    preserving rank order.  Returns the concatenated list."""
    import torch
    world = dist.get_world_size()
    blob = '\x1e'.join(strings).encode()
    local = torch.tensor([len(blob), len(strings)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, local)
    maxlen = max(int(s[0].item()) for s in sizes)
    buf = torch.zeros(max(1, maxlen), dtype=torch.uint8, device=device)
    if blob:
        buf[:len(blob)] = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).to(device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    out = []
    for s, b in zip(sizes, bufs):
        n, cnt = int(s[0].item()), int(s[1].item())
        if cnt == 0:
            continue
        out.extend(bytes(b[:n].cpu().numpy().tobytes()).decode().split('\x1e'))
    return out
