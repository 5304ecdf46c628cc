"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: cost-balanced read partition, reference broadcast
and the variable-length result all-gather used by bench.py --gpus N."""
import os
import subprocess
import sys
import textwrap

from oracle_lib import ROOT


def _free_port():
    import socket
    with socket.socket() as sock:
        sock.bind(('127.0.0.1', 0))
        return sock.getsockname()[1]


def test_partition_by_cost_balances_and_covers():
    sys.path.insert(0, ROOT)
    from unicycler_b200.sharding import partition_by_cost
    costs = [5, 1, 1, 1, 8, 2, 2, 9, 3, 1, 1, 7]
    for world in (1, 2, 3, 4, 8, 16):
        parts = partition_by_cost(costs, world)
        assert len(parts) == world
        assert parts[0][0] == 0 and parts[-1][1] == len(costs)
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and a <= b
        if world == 2:
            s0 = sum(costs[parts[0][0]:parts[0][1]])
            assert abs(s0 - sum(costs) / 2) <= max(costs)


def test_broadcast_and_all_gather_world2(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(textwrap.dedent('''
        import os, sys
        sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        from unicycler_b200 import sharding
        rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
        dist.init_process_group('gloo', rank=rank, world_size=world)
        dev = torch.device('cpu')
        refs = [('chr1', 'ACGT' * 50), ('plasmid', 'TTGACA' * 7)] if rank == 0 else None
        got = sharding.broadcast_references(refs, dist, dev)
        assert got == [('chr1', 'ACGT' * 50), ('plasmid', 'TTGACA' * 7)], got
        reads = ['r%%d' %% i for i in range(7)]
        costs = [3, 1, 4, 1, 5, 9, 2]
        a, b = sharding.partition_by_cost(costs, world)[rank]
        mine = ['%%s:aligned-by-%%d;' %% (r, rank) * (1 + i %% 3) for i, r in enumerate(reads[a:b])]
        allres = sharding.all_gather_strings(mine, dist, dev)
        assert len(allres) == len(reads), (len(allres), allres)
        assert [x.split(':')[0] for x in allres] == reads
        if rank == 1:
            empty = sharding.all_gather_strings([], dist, dev)
        else:
            empty = sharding.all_gather_strings(['x'], dist, dev)
        assert empty == ['x']
        dist.destroy_process_group()
        sys.stdout.write('rank {} ok'.format(rank) + chr(10))
        sys.stdout.flush()
    ''' % ROOT))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29541')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), str(script)],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=240)
    text = out.stdout.decode()
    assert out.returncode == 0, text
    assert 'rank 0 ok' in text and 'rank 1 ok' in text, text
