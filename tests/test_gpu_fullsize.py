"""BASELINE.json-size batches (4096 windows, 16384 LS epochs, 256 x 64 EKF steps) checked through
size-independent properties: every window converges, re-solving from the solution is a fixed
point, results do not depend on the window's position in the batch, the gradient vanishes at
the returned point; plus oracle parity on a random subset."""
import numpy as np
import pytest

from nlp_filter_b200 import desc as D
from nlp_filter_b200 import shapes, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload", ["vdp", "gnss"])
def test_4096_windows_properties(workload):
    import torch
    from nlp_filter_b200 import engine
    from oracle import nls
    B = 4096
    d = shapes.vdp_mhe() if workload == "vdp" else shapes.gnss()
    bt = synth.vdp_mhe_batch(d, B, seed=2000) if workload == "vdp" else synth.gnss_batch(d, B, seed=3000)
    prob = engine.WindowProblem(d)
    z, J, iters, status = engine.solve_windows_host(prob, bt["z0"], bt["U"], bt["data"])
    assert np.all(status <= 1), np.bincount(status)                 # every window solved
    assert np.all(np.isfinite(z)) and np.max(iters) <= 60
    # fixed point: restarting at the solution moves nothing
    z2, J2, it2, st2 = engine.solve_windows_host(prob, z, bt["U"], bt["data"])
    assert np.max(np.abs(z2 - z) / (1 + np.abs(z))) < 1e-8 and np.max(it2) <= 6
    # stationarity: the gradient at the solution is at rounding level relative to its start value
    _, g0, _ = prob.assemble(bt["z0"][:256], bt["U"][:256], bt["data"][:256])
    _, g1, _ = prob.assemble(z[:256], bt["U"][:256], bt["data"][:256])
    assert float(g1.abs().max()) < 1e-7 * float(g0.abs().max())
    # position independence: a permuted batch gives bit-identical per-window results
    perm = np.random.default_rng(0).permutation(B)
    zp, Jp, _, _ = engine.solve_windows_host(prob, bt["z0"][perm], bt["U"][perm], bt["data"][perm])
    assert np.array_equal(zp, z[perm]) and np.array_equal(Jp, J[perm])
    # oracle parity on a random subset
    for b in np.random.default_rng(1).choice(B, 12, replace=False):
        zo, info = nls.solve(d, nls.Window(d, U=bt["U"][b], data=bt["data"][b]), bt["z0"][b])
        assert np.max(np.abs(z[b] - zo) / (1 + np.abs(zo))) < 1e-6
        if workload == "gnss":
            assert np.max(np.abs(z[b].reshape(d.K, 5)[:, :3] - zo.reshape(d.K, 5)[:, :3])) < 1e-3     # 1 mm
    prob.close()


def test_16384_ls_epochs_properties():
    """LS: solution is a fixed point of one more Gauss-Newton step and is translation-covariant."""
    from nlp_filter_b200 import engine
    from oracle import ekf_ls
    rng = np.random.default_rng(5)
    B, S = 16384, 12
    p_ref = np.array([-2699899.3380952473, -4292848.690053425, 3855186.0888884724])
    rx = p_ref[None, :] + rng.normal(0, 50.0, (B, 3))
    dirs = rng.normal(size=(B, S, 3)) + 1.5 * (p_ref / np.linalg.norm(p_ref))[None, None, :]
    dirs /= np.linalg.norm(dirs, axis=2, keepdims=True)
    sat = rx[:, None, :] + dirs * rng.uniform(2.0e7, 2.5e7, (B, S, 1))
    bias = rng.uniform(-1e3, 1e3, B)
    pr = np.linalg.norm(sat - rx[:, None, :], axis=2) + bias[:, None] + rng.normal(0, 3.0, (B, S))
    nsat = rng.integers(6, S + 1, B).astype(np.int32)
    x, it = engine.ls_solve(sat, pr, nsat)
    x, it = x.cpu().numpy(), it.cpu().numpy()
    assert np.all(it <= 20) and np.all(np.isfinite(x)), (it.max(), np.isfinite(x).all())
    assert np.median(np.linalg.norm(x[:, :3] - rx, axis=1)) < 15.0      # metres: 3 m range noise times the geometry factor
    x2, it2 = engine.ls_solve(sat, pr, nsat, x0=x)                    # warm restart: converged already
    it2 = it2.cpu().numpy()
    # the stopping rule |dx|_2 < 1e-7 m (leastsquares.py:39) sits at the rounding level of the pseudorange residuals for
    # the few epochs with poor geometry (6 satellites): those need some more passes to get under it, as in the reference
    assert np.max(np.abs(x2.cpu().numpy() - x)) < 1e-6 and np.mean(it2 <= 3) > 0.999 and it2.max() <= 30, (np.mean(it2 <= 3), it2.max())
    shift = np.array([1000.0, -2000.0, 500.0])                        # translate the whole geometry
    xs, _ = engine.ls_solve(sat + shift, pr, nsat)
    assert np.max(np.abs(xs.cpu().numpy()[:, :3] - (x[:, :3] + shift))) < 1e-5
    for b in rng.choice(B, 16, replace=False):
        xo, bo, _ = ekf_ls.iterative_ls(sat[b, :nsat[b]], pr[b, :nsat[b]])
        assert np.max(np.abs(x[b] - np.append(xo, bo)) / (1 + np.abs(x[b]))) < 1e-9


def test_256_filters_64_steps_against_oracle_subset():
    from nlp_filter_b200 import engine
    from oracle import ekf_ls
    rng = np.random.default_rng(6)
    Bf, steps, S = 256, 64, 12
    Q = np.array([0.0001, 0.0001, 0.0001, 0.1, 0.001])
    pos = rng.normal(0, 5.0, (Bf, 3))
    b0, al = rng.uniform(-1e3, 1e3, Bf), rng.normal(-190, 20, Bf)
    el = np.arcsin(rng.uniform(np.sin(np.deg2rad(10)), 1.0, (Bf, S)))
    az = rng.uniform(0, 2 * np.pi, (Bf, S))
    rad = rng.uniform(2.0e7, 2.6e7, (Bf, S))
    s0 = np.stack([rad * np.cos(el) * np.sin(az), rad * np.cos(el) * np.cos(az), rad * np.sin(el)], 2)
    t = np.arange(1, steps + 1)
    sat = s0[:, None, :, :] + rng.normal(0, 1e3, (Bf, 1, S, 3)) * t[None, :, None, None]
    z = np.linalg.norm(pos[:, None, None, :] - sat, axis=3) + b0[:, None, None] + al[:, None, None] * t[None, :, None] \
        + rng.normal(0, 10.0, (Bf, steps, S))
    nsat = rng.integers(4, S + 1, (Bf, steps)).astype(np.int32)
    nsat[:, ::7] = 0                                                  # predict-only steps (ekf.py:32-38)
    x0 = np.concatenate([pos + rng.normal(0, 3, (Bf, 3)), (b0 + 5)[:, None], (al + 1)[:, None]], 1)
    u = np.zeros((Bf, steps, 3))
    mu, S_, traj = engine.ekf_run(engine.EKF_GNSS5, x0, np.tile(np.eye(5), (Bf, 1, 1)), u, z, sat, nsat, Q, 100.0, 1.0)
    traj = traj.cpu().numpy()
    assert np.all(np.isfinite(traj))
    assert np.median(np.linalg.norm(traj[:, -1, :3] - pos, axis=1)) < 5.0
    for b in rng.choice(Bf, 8, replace=False):
        ref, Sref = ekf_ls.ekf_run(0, x0[b], np.eye(5), u[b], z[b], sat[b], nsat[b], Q, 100.0, 1.0)
        assert np.max(np.abs(traj[b] - ref) / (1 + np.abs(ref))) < 1e-9
        assert np.max(np.abs(S_.cpu().numpy()[b] - Sref)) < 1e-9


@pytest.mark.parametrize("mu0_warm", [0.0, 1e-6])
def test_vehicle_chains_against_oracle_and_truth(mu0_warm):
    """Synthetic receding-horizon vehicle chains (autonomous-car.py:232-288 pattern): device-side chain
    driver vs the oracle run through the same chain, and estimates close to the simulated truth.  mu0_warm: the
    warm-started windows 2.. start the barrier there (nlpf_opts.mu0_warm; 0 = at mu0 like the first window)."""
    import torch
    from nlp_filter_b200 import engine
    from oracle import nls
    d = shapes.car()
    opts = D.SolveOpts(mu0_warm=mu0_warm)
    B, W = 64, 5
    ch = synth.car_chain_batch(d, B, W, seed=11)
    prob = engine.WindowProblem(d)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    z = torch.zeros((B, d.nz), dtype=torch.float64, device="cuda")
    xp = dev(ch["xhat0"])
    it_sum, bad, Z = engine.solve_chain(prob, z, dev(ch["U"]), dev(ch["data"]), xp, ch["PhiDT"], opts=opts, keep=True)
    torch.cuda.synchronize()
    assert int(bad.sum()) == 0
    if mu0_warm > 0.0:
        assert float(it_sum.sum()) / (B * W) < 12.0          # (the full barrier schedule per window: 14-15)
    Z = Z.cpu().numpy()
    V = Z[-1].reshape(B, d.K, d.nv)
    assert np.median(np.linalg.norm(V[:, 0, :2] - ch["truth"][-1][:, :2], axis=1)) < 1.5      # metres, GNSS noise level
    for b in (0, 17, 63):
        zo = np.zeros(d.nz)
        xpo = ch["xhat0"][b].copy()
        for w in range(W):
            zo, info = nls.solve(d, nls.Window(d, U=ch["U"][w, b], data=ch["data"][w, b], xprior=xpo), zo, opts.for_window(w))
            xpo = (ch["PhiDT"] @ zo.reshape(d.K, d.nv))[0]
            Vg, Vo = Z[w, b].reshape(d.K, d.nv), zo.reshape(d.K, d.nv)
            assert np.max(np.abs(Vg[:, [0, 1, 8]] - Vo[:, [0, 1, 8]])) < 1e-3                  # positions within 1 mm
            assert np.max(np.abs(Vg - Vo) / (1 + np.abs(Vo))) < 1e-5
    prob.close()


@pytest.mark.parametrize("mu0_warm", [0.0, 1e-6])
def test_fused_chain_kernel_equals_the_per_window_loop(monkeypatch, mu0_warm):
    """nlpf_solve_chain (every run's windows back to back in one warp) against the host loop of
    nlpf_solve + nlpf_sample launches it replaces: same kernel code per window, so identical results."""
    import os
    import torch
    from nlp_filter_b200 import engine
    if os.environ.get("NLPF_NO_WARP_KERNEL"):
        pytest.skip("the fused chain lives in the warp-per-window kernel, which this run disables")
    for loss in (D.LOSS_L2, D.LOSS_HUBER):
        d = shapes.car(loss=loss)
        B, W = 96, 4
        ch = synth.car_chain_batch(d, B, W, seed=23)
        prob = engine.WindowProblem(d)
        assert prob._lib.nlpf_problem_has_chain_kernel(prob._h) == 1
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
        out = []
        for loop in (False, True):
            if loop:
                monkeypatch.setenv("NLPF_CHAIN_LOOP", "1")
            else:
                monkeypatch.delenv("NLPF_CHAIN_LOOP", raising=False)
            z = torch.zeros((B, d.nz), dtype=torch.float64, device="cuda")
            xp = dev(ch["xhat0"])
            it_sum, bad, Z = engine.solve_chain(prob, z, dev(ch["U"]), dev(ch["data"]), xp, ch["PhiDT"],
                                                opts=D.SolveOpts(mu0_warm=mu0_warm), keep=True)
            torch.cuda.synchronize()
            out.append((it_sum.cpu().numpy(), bad.cpu().numpy(), Z.cpu().numpy(), z.cpu().numpy(), xp.cpu().numpy()))
        f, l = out
        assert np.array_equal(f[0], l[0]) and np.array_equal(f[1], l[1])
        assert np.array_equal(f[2], l[2]) and np.array_equal(f[3], l[3])
        assert np.max(np.abs(f[4] - l[4])) <= 1e-12 * (1 + np.max(np.abs(l[4])))      # sampled by two different kernels
        prob.close()
    # a class without arrival cost has no chain kernel: the library says so instead of guessing
    pv = engine.WindowProblem(shapes.vdp_mhe())
    assert pv._lib.nlpf_problem_has_chain_kernel(pv._h) == 0
    pv.close()


def test_host_chains_equal_the_device_resident_chain_call():
    """engine.HostChains (bench.py's end-to-end leg of the vehicle workload: groups of runs with their own streams,
    handles and run-major pinned buffers) against one solve_chain call on device-resident inputs: runs are
    independent, so grouping must not change a bit."""
    import torch
    from nlp_filter_b200 import engine
    d = shapes.car()
    B, W = 50, 4
    ch = synth.car_chain_batch(d, B, W, seed=29)
    opts = D.SolveOpts(mu0_warm=1e-6)
    prob = engine.WindowProblem(d)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    z = torch.zeros((B, d.nz), dtype=torch.float64, device="cuda")
    xp = dev(ch["xhat0"])
    it_ref, bad_ref = engine.solve_chain(prob, z, dev(ch["U"]), dev(ch["data"]), xp, ch["PhiDT"], opts=opts)
    torch.cuda.synchronize()
    for nchunk in (1, 3, 7):
        hc = engine.HostChains(prob, B, W, nchunk=nchunk)
        hc.load(ch["U"], ch["data"], ch["xhat0"])
        hc.run(ch["PhiDT"], opts)
        hc.run(ch["PhiDT"], opts)                      # staged inputs survive a run
        torch.cuda.synchronize()
        zz, it, bad = hc.results()
        assert np.array_equal(zz, z.cpu().numpy())
        assert np.array_equal(it, it_ref.cpu().numpy()) and np.array_equal(bad, bad_ref.cpu().numpy())
        hc.close()
    prob.close()


def test_chunked_end_to_end_batches_equal_the_single_chunk_run():
    """HostBatch cuts a batch into chunks with their own streams, handles and packed staging buffers (what
    bench.py's e2e leg uses at 4096 windows); chunking must not change a bit of the results."""
    import torch
    from nlp_filter_b200 import engine
    for d, bt in ((shapes.vdp_mhe(), None), (shapes.car(), None)):
        B = 50
        if d.has_prior:
            ch = synth.car_chain_batch(d, B, 1, seed=41)
            z0, U, data, xp = np.zeros((B, d.nz)), ch["U"][0], ch["data"][0], ch["xhat0"]
        else:
            b2 = synth.vdp_mhe_batch(d, B, seed=42)
            z0, U, data, xp = b2["z0"], b2["U"], b2["data"], None
        prob = engine.WindowProblem(d)
        outs = []
        for nchunk in (1, 3, 7):
            hb = engine.HostBatch(prob, B, nchunk=nchunk)
            assert hb.nchunk == nchunk
            hb.load(z0, U, data, xp)
            hb.run()
            hb.run()                                   # staged inputs survive a run
            torch.cuda.synchronize()
            outs.append([a.copy() for a in hb.results()] + [hb.device_z().cpu().numpy()])
        # the same step recorded as a CUDA graph (HostBatch.capture) and replayed: same bits, also after new inputs
        hb = engine.HostBatch(prob, B, nchunk=3)
        hb.load(z0, U, data, xp)
        hb.run()
        hb.capture()
        hb.run()
        torch.cuda.synchronize()
        outs.append([a.copy() for a in hb.results()] + [hb.device_z().cpu().numpy()])
        for o in outs[1:]:
            for a, b in zip(o, outs[0]):
                assert np.array_equal(a, b)
        hb.load(z0[::-1].copy(), None if U is None else U[::-1].copy(), data[::-1].copy(), None if xp is None else xp[::-1].copy())
        hb.run()
        torch.cuda.synchronize()
        assert np.array_equal(hb.results()[0], outs[0][0][::-1])
        assert np.array_equal(outs[0][0], outs[0][4]) and np.all(outs[0][3] <= 1)
        prob.close()


def test_2048_vehicle_runs_of_200_windows():
    """BASELINE configs[3] at its stated size: 2048 runs x 200 chained windows (409 600 solves) in one launch.
    Size-independent properties plus the oracle on a random subset of windows ANYWHERE in the chains: a window of
    a chain depends only on its predecessor's estimate, so the oracle is started from the device's own
    predecessor (warm start + re-centred arrival cost) and must land on the device's estimate.  Windows that end
    at the iteration limit are located, counted and compared with what the oracle does from the same start."""
    import torch
    from nlp_filter_b200 import engine
    from oracle import nls
    d = shapes.car()
    B, W = 2048, 200
    ch = synth.car_chain_batch(d, B, W, seed=4000)
    prob = engine.WindowProblem(d)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    z = torch.zeros((B, d.nz), dtype=torch.float64, device="cuda")
    xp = dev(ch["xhat0"])
    U_d, data_d = dev(ch["U"]), dev(ch["data"])
    opts = D.SolveOpts(mu0_warm=1e-6)                            # what bench.py --workload car runs with
    it_sum, bad, Z = engine.solve_chain(prob, z, U_d, data_d, xp, ch["PhiDT"], opts=opts, keep=True)
    torch.cuda.synchronize()
    Z = Z.cpu().numpy()
    bad = bad.cpu().numpy()
    assert np.all(np.isfinite(Z))
    n_bad = int(bad.sum())
    assert n_bad <= 41, n_bad                                   # <= 1e-4 of the 409 600 windows (measured: see DESIGN)
    # estimates follow the simulated truth over the whole run (GNSS noise level), bad windows included
    V0 = Z.reshape(W, B, d.K, d.nv)[:, :, 0, :]
    err = np.linalg.norm(V0[:, :, :2] - ch["truth"][:, :, :2], axis=2)
    assert np.median(err) < 1.5 and np.max(err[20:]) < 50.0, (np.median(err), np.max(err[20:]))
    PhiDT = ch["PhiDT"]

    def oracle_window(w, b):
        zprev = Z[w - 1, b] if w > 0 else np.zeros(d.nz)
        xpo = (PhiDT @ zprev.reshape(d.K, d.nv))[0] if w > 0 else ch["xhat0"][b]
        return nls.solve(d, nls.Window(d, U=ch["U"][w, b], data=ch["data"][w, b], xprior=xpo), zprev, opts.for_window(w))

    rng = np.random.default_rng(3)
    for w, b in zip(rng.integers(0, W, 10), rng.integers(0, B, 10)):
        zo, info = oracle_window(int(w), int(b))
        Vg, Vo = Z[w, b].reshape(d.K, d.nv), zo.reshape(d.K, d.nv)
        assert np.max(np.abs(Vg[:, [0, 1, 8]] - Vo[:, [0, 1, 8]])) < 1e-3, (w, b)             # positions within 1 mm
        assert np.max(np.abs(Vg - Vo) / (1 + np.abs(Vo))) < 1e-5, (w, b)
    # the windows that did not converge: find them by re-solving the bad runs window by window from the stored
    # predecessors (same kernel code, so the same status), then ask the oracle about the same windows
    checked = 0
    for b in np.nonzero(bad)[0][:4]:
        zs = np.concatenate([np.zeros((1, d.nz)), Z[:-1, b]], 0)
        xps = np.concatenate([ch["xhat0"][b][None], (zs[1:].reshape(W - 1, d.K, d.nv)[:, :, :] * PhiDT[0][None, :, None]).sum(1)], 0)
        # (windows 1.. in one batch with the warm-start options; window 0 of a run, the cold one, is not replayed)
        zz, J, iters, status = engine.solve_windows_host(prob, zs[1:], ch["U"][1:, b], ch["data"][1:, b], xps[1:], opts=opts.for_window(1))
        zz = np.concatenate([Z[:1, b], zz], 0)
        iters, status = np.concatenate([[0], iters]), np.concatenate([[0], status])
        ws = np.nonzero(status > 1)[0]
        good = status <= 1
        # replay of the converged windows.  The arrival cost is re-sampled on the host (differences of 1e-12); most
        # windows then reproduce to 1e-9, but the slowly converging ones (60+ Gauss-Newton iterations: the vehicle
        # model carries no second-order term) sit in a flat valley where the step criterion leaves the velocity
        # components undetermined at the 1e-3 level -- the oracle, started from the same point, differs from the
        # device by the same amount (measured: positions <= 4.3e-4 m, vy <= 3.6e-3 m/s, J equal to 1e-9 relative).
        # Asserted: positions and clock bias within 1 mm, everything else within 1e-2.
        dV = np.abs(zz[good] - Z[good, b]).reshape(-1, d.K, d.nv)
        assert np.max(dV[:, :, [0, 1, 8, 6]]) < 1e-3, np.max(dV, (0, 1))
        assert np.max(dV) < 1e-2, np.max(dV, (0, 1))
        print("run %d: windows at the limit %s (chain kernel counted %d), iterations there %s" % (b, ws.tolist(), bad[b], iters[ws].tolist()))
        for w in ws[:2]:
            zo, info = oracle_window(int(w), int(b))
            # the oracle, started from the same point, does not converge either (status 2); the two unconverged
            # iterates are compared at a centimetre (they wander along the same flat valley, each with its own
            # rounding), and the chain recovers: the next window converges
            Vg, Vo = Z[w, b].reshape(d.K, d.nv), zo.reshape(d.K, d.nv)
            print("   window %d: oracle %s; |dpos| = %.3e m, max rel %.3e" % (w, info, np.max(np.abs(Vg[:, [0, 1, 8]] - Vo[:, [0, 1, 8]])),
                                                                     np.max(np.abs(Vg - Vo) / (1 + np.abs(Vo)))))
            assert np.max(np.abs(Vg[:, [0, 1, 8]] - Vo[:, [0, 1, 8]])) < (1e-3 if info["status"] <= 1 else 1e-2), (w, b, status[w], info)
            if w + 1 < W:
                assert status[w + 1] <= 1
            checked += 1
    print("vehicle 2048 x 200: %d windows at the iteration limit, %d checked against the oracle; mean %.1f iterations"
          % (n_bad, checked, float(it_sum.sum()) / (B * W)))
    prob.close()


def test_1024_two_receiver_scenarios():
    """BASELINE configs[4] at its stated size: 1024 two-receiver scenarios in one launch; every scenario
    converges, re-solving is a fixed point, the folded equality z_A = z_B holds, oracle parity on a subset."""
    from nlp_filter_b200 import engine
    from oracle import nls
    d = shapes.two_receiver()
    B = 1024
    bt = synth.two_receiver_batch(d, B, seed=5000)
    prob = engine.WindowProblem(d)
    z, J, iters, status = engine.solve_windows_host(prob, bt["z0"], bt["U"], bt["data"], bt["xprior"])
    assert np.all(status <= 1), np.bincount(status)
    assert np.all(np.isfinite(z))
    z2, J2, it2, st2 = engine.solve_windows_host(prob, z, bt["U"], bt["data"], bt["xprior"])
    assert np.max(np.abs(z2 - z) / (1 + np.abs(z))) < 1e-7 and np.all(st2 <= 1)
    perm = np.random.default_rng(0).permutation(B)
    zp, Jp, _, _ = engine.solve_windows_host(prob, bt["z0"][perm], bt["U"][perm], bt["data"][perm], bt["xprior"][perm])
    assert np.array_equal(zp, z[perm])
    for b in np.random.default_rng(1).choice(B, 8, replace=False):
        zo, info = nls.solve(d, nls.Window(d, U=bt["U"][b], data=bt["data"][b], xprior=bt["xprior"][b]), bt["z0"][b])
        V, Vo = z[b].reshape(d.K, d.nv), zo.reshape(d.K, d.nv)
        assert np.max(np.abs(V - Vo) / (1 + np.abs(Vo))) < 1e-6, b
        assert np.max(np.abs(V[:, :3] - Vo[:, :3])) < 1e-3                                     # receiver A position, 1 mm
    prob.close()
