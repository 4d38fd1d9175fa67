"""Launched under torchrun by tests/test_gpu_multi.py (one rank per GPU, NCCL).

Every rank builds its z-slab of the same box mesh and runs the device-resident LVPP loop; rank 0
also solves the whole mesh on its own GPU and checks: residual, J*v and observables of the
partitioned problem against the single-GPU ones on the same global state, then Newton counts per
proximal step and the final solution."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import proximalgalerkin_b200 as lvpp  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    shape = tuple(int(v) for v in sys.argv[1].split("x"))
    make = lvpp.mesh.create_box if len(shape) == 3 else lvpp.mesh.create_rectangle
    msh = make(*shape, rank=rank, nranks=world)
    whole = make(*shape)
    popts = {"ksp_type": "gmres", "pc_type": "mg"} if len(sys.argv) > 2 and sys.argv[2] == "mg" else None
    s = lvpp.obstacle_pg.setup(msh, 1, petsc_options=popts)
    dev = s["problem"].device_problem
    nown = msh.num_owned_vertices
    gv = msh.global_vertex

    # ---- same global state on both layouts
    rng = np.random.default_rng(5)
    xg = 0.3 * rng.standard_normal(2 * whole.num_vertices)
    xkg = 0.3 * rng.standard_normal(2 * whole.num_vertices)

    def local(vec):  # global mixed vector -> local owned part (ghosts poisoned: the library must refresh them)
        out = np.full(2 * msh.num_vertices, 1e300)
        out[0 : 2 * nown : 2] = vec[2 * gv[:nown]]
        out[1 : 2 * nown : 2] = vec[2 * gv[:nown] + 1]
        return out

    alpha = 2.5
    dev.set_alpha(alpha)
    dev.set_previous(local(xkg))
    X, F, Y = (lvpp.DeviceVector(dev.n, dev.device) for _ in range(3))
    X.set(local(xg))
    fnorm = dev.assemble_residual(X, F)
    dev.spmv(X, Y)
    obs = dev.observables(X)
    ok = True
    if rank == 0:
        s1 = lvpp.obstacle_pg.setup(whole, 1, petsc_options=popts)
        d1 = s1["problem"].device_problem
        d1.set_alpha(alpha)
        d1.set_previous(xkg)
        X1, F1, Y1 = (lvpp.DeviceVector(d1.n, d1.device) for _ in range(3))
        X1.set(xg)
        fnorm1 = d1.assemble_residual(X1, F1)
        d1.spmv(X1, Y1)
        obs1 = d1.observables(X1)
        ref = {"F": F1.numpy(), "Y": Y1.numpy()}
        assert abs(fnorm - fnorm1) <= 1e-12 * fnorm1, (fnorm, fnorm1)
        assert np.allclose(obs, obs1, rtol=1e-12, atol=1e-14), (obs, obs1)
        assert dev.stats()["num_rows"] == 2 * whole.num_vertices
        torch.save(ref, "/tmp/lvpp_multi_ref.pt")
    dist.barrier()
    ref = torch.load("/tmp/lvpp_multi_ref.pt", weights_only=False)
    for name, vec in (("F", F.numpy()), ("Y", Y.numpy())):
        mine = vec[: 2 * nown].reshape(-1, 2)
        want = ref[name].reshape(-1, 2)[gv[:nown]]
        err = np.abs(mine - want).max() / np.abs(want).max()
        assert err < 1e-13, (name, rank, err)

    # ---- full LVPP solve, partitioned vs single GPU
    st = lvpp.obstacle_pg.LvppStepper(msh, 1, "double_exponential", 1e2, 1e-4, setup_objects=s)
    while st.step():
        pass
    if rank == 0:
        st1 = lvpp.obstacle_pg.LvppStepper(whole, 1, "double_exponential", 1e2, 1e-4, setup_objects=s1)
        while st1.step():
            pass
        assert st.history["newton_steps"] == st1.history["newton_steps"], (st.history, st1.history)
        print(f"single-GPU krylov={d1.stats()['krylov_iterations']} partitioned krylov={dev.stats()['krylov_iterations']}")
        torch.save({"x": st1.x.numpy()}, "/tmp/lvpp_multi_sol.pt")
    dist.barrier()
    sol = torch.load("/tmp/lvpp_multi_sol.pt", weights_only=False)["x"].reshape(-1, 2)
    mine = st.x.numpy()[: 2 * nown].reshape(-1, 2)
    err = np.linalg.norm(mine[:, 0] - sol[gv[:nown], 0]) / np.linalg.norm(sol[:, 0])
    assert err < 1e-10, (rank, err)
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} shape={shape} pc={'mg' if popts else 'jacobi'} krylov={dev.stats()['krylov_iterations']} newton={st.history['newton_steps']} u_err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
