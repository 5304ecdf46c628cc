"""Multi-GPU plumbing for the alignment path: one process per GPU (torch.distributed), reads partitioned across
ranks by estimated DP cost, reference set broadcast once, variable-length result strings gathered with exact sizes.  No collective sits inside the DP: every (read, reference range) job is independent."""
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


def gather_strings(strings, dist, device, dst=0):
    """Gathers each rank's list of result strings on rank `dst` in rank order (None elsewhere) with exact sizes: one
    all-gather of the byte counts, then one grouped send/recv of exactly those bytes (ncclSend/ncclRecv inside one
    group on the GPU box, gloo in the CPU tests).  Results end on the host of the gathering rank anyway, so nothing
    is padded and nothing is sent to ranks that do not need it."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    blob = '\x1e'.join(strings).encode()
    local = torch.tensor([len(blob), len(strings)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, local)
    sizes = [(int(s[0].item()), int(s[1].item())) for s in sizes]
    ops, bufs = [], {}
    if rank == dst:
        for r, (n, cnt) in enumerate(sizes):
            if r != dst and n > 0:
                bufs[r] = torch.empty(n, dtype=torch.uint8, device=device)
                ops.append(dist.P2POp(dist.irecv, bufs[r], r))
    elif blob:
        mine = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).to(device)
        ops.append(dist.P2POp(dist.isend, mine, dst))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    out = []
    for r, (n, cnt) in enumerate(sizes):
        if cnt == 0:
            continue
        data = blob if r == dst else bytes(bufs[r].cpu().numpy().tobytes())
        out.extend(data.decode().split('\x1e'))
    return out


def all_gather_strings(strings, dist, device):
    """Every rank gets the concatenated list (rank order): gather on rank 0 with exact sizes, then one broadcast of
    the packed bytes."""
    import torch
    rank = dist.get_rank()
    merged = gather_strings(strings, dist, device, dst=0)
    blob = '\x1e'.join(merged).encode() if rank == 0 else b''
    meta = torch.tensor([len(blob), len(merged) if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(meta, src=0)
    n, cnt = int(meta[0].item()), int(meta[1].item())
    if cnt == 0:
        return []
    if rank == 0:
        buf = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).to(device)
    else:
        buf = torch.empty(n, dtype=torch.uint8, device=device)
    if n > 0:
        dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes()).decode().split('\x1e') if n > 0 else [''] * cnt
