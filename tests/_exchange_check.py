"""N-rank check of the interface exchange (test infrastructure; launched under torchrun by
tests/test_gpu_cg.py::test_exchange_multi_gpu and by bench.py's sharded leg).

Random universal-id maps in which a DOF is held by 1..N ranks (the general gslib situation: edges and corners of a
box partition, AssemblyMapCG.cpp:2551-2569) -> peers / ordered lists / ownership mask from
mesh.interface_from_universal_maps -> nekmf_exchange_add on the device.  Expected values: for every universal id the
sum of the holders' values in ascending rank order, computed in numpy from the all-gathered inputs -- the device
result must be BIT-IDENTICAL to it on every holder (deterministic rank-ordered unpack), for several consecutive
exchanges (parity double buffering) and for both transports."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from _util import load_pkg_module, nekmf  # noqa: E402


def main():
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nk = nekmf()
    mesh_mod = load_pkg_module("mesh")
    comm = nk.Comm.from_torch_distributed()
    rng = np.random.default_rng(99)  # same stream on every rank: everybody knows every map
    nU = 20000
    maps = []
    for r in range(world):
        held = np.flatnonzero(rng.random(nU) < 0.6) + 1       # universal ids held by rank r (0 = not taking part)
        extra = np.zeros(500, dtype=np.int64)                 # some DOFs of the rank do not take part
        m = np.concatenate([held, extra])
        rng.shuffle(m)
        maps.append(m)
    mine = maps[rank]
    nG = mine.size
    peers, lists, owner = mesh_mod.interface_from_universal_maps(maps, rank)
    ex = nk.Exchange(comm, peers, lists, nG)
    ok = True
    for rep in range(5):
        vals = [np.random.default_rng(1000 * rep + r).uniform(-1, 1, maps[r].size) for r in range(world)]
        want = vals[rank].copy()
        # rank-ordered sum per universal id
        total = np.zeros(nU + 1)
        started = np.zeros(nU + 1, dtype=bool)
        for r in range(world):
            ids = maps[r]
            sel = ids != 0
            first = sel & ~started[ids]
            total[ids[first]] = vals[r][first]
            later = sel & started[ids]
            total[ids[later]] = total[ids[later]] + vals[r][later]
            started[ids[sel]] = True
        sel = mine != 0
        want[sel] = total[mine[sel]]
        g = torch.tensor(vals[rank], device=dev)
        ex.add(g)
        torch.cuda.synchronize()
        got = g.cpu().numpy()
        if not np.array_equal(got, want):
            ok = False
            print("rank %d rep %d: max diff %.3e (%d entries differ)" % (
                rank, rep, np.abs(got - want).max(), int((got != want).sum())))
    # the ownership mask counts every universal id once
    cnt = torch.tensor([float(owner[mine != 0].sum())], device=dev)
    dist.all_reduce(cnt)
    n_ids = len(set(np.concatenate([m[m != 0] for m in maps]).tolist()))
    if int(cnt.item()) != n_ids:
        ok = False
        print("ownership mask counts %d ids, expected %d" % (int(cnt.item()), n_ids))
    t = torch.tensor([0.0 if ok else 1.0], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("exchange transport=%s ranks=%d peers(rank0)=%s" % (comm.transport, world, peers))
        print("CHECK OK" if float(t.item()) == 0.0 else "CHECK FAILED")
    del ex
    dist.barrier()
    del comm
    dist.destroy_process_group()
    sys.exit(0 if float(t.item()) == 0.0 else 1)


if __name__ == "__main__":
    main()
