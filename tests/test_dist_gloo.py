"""World-size-2 run of the host-side multi-GPU logic on CPU (gloo): each rank builds its slab,
ghost values are exchanged along the halo lists exactly as lvpp_halo_forward does on the device
(pack owned -> send/recv -> unpack into ghosts), and a distributed dot product is all-reduced."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape):
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import proximalgalerkin_b200 as lvpp

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        make = lvpp.mesh.create_box if len(shape) == 3 else lvpp.mesh.create_rectangle
        msh = make(*shape, rank=rank, nranks=world)
        whole = make(*shape)
        # mixed vector: owned entries = function of the global vertex number, ghosts poisoned
        v = np.full((msh.num_vertices, 2), np.nan)
        gv = msh.global_vertex[: msh.num_owned_vertices]
        v[: msh.num_owned_vertices, 0] = np.sin(gv)
        v[: msh.num_owned_vertices, 1] = np.cos(gv)
        reqs, bufs = [], []
        for nb, send, recv in zip(msh.halo.neighbors, msh.halo.send, msh.halo.recv):
            sb = torch.from_numpy(np.ascontiguousarray(v[send]))
            rb = torch.empty((recv.size, 2), dtype=torch.float64)
            reqs.append(dist.isend(sb, nb))
            reqs.append(dist.irecv(rb, nb))
            bufs.append((recv, rb))
        for r in reqs:
            r.wait()
        for recv, rb in bufs:
            v[recv] = rb.numpy()
        g = msh.global_vertex
        assert np.array_equal(v[:, 0], np.sin(g)) and np.array_equal(v[:, 1], np.cos(g))
        # distributed dot over owned rows == global dot
        part = torch.tensor([np.sum(v[: msh.num_owned_vertices] ** 2)], dtype=torch.float64)
        dist.all_reduce(part)
        gg = np.arange(whole.num_vertices)
        assert abs(part.item() - np.sum(np.sin(gg) ** 2 + np.cos(gg) ** 2)) < 1e-9
        # the NCCL unique id travels as a 128-byte broadcast (DeviceProblem._init_comm)
        t = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(t, src=0)
        assert t.tolist() == list(range(128))
        # global row count the library would all-reduce
        rows = torch.tensor([2.0 * msh.num_owned_vertices], dtype=torch.float64)
        dist.all_reduce(rows)
        assert int(rows.item()) == 2 * whole.num_vertices
    finally:
        dist.destroy_process_group()


def test_halo_exchange_world2_box():
    mp.spawn(_worker, args=(2, _free_port(), (3, 4, 6)), nprocs=2, join=True)


def test_halo_exchange_world2_rectangle():
    mp.spawn(_worker, args=(2, _free_port(), (5, 8)), nprocs=2, join=True)
