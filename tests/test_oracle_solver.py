"""Main-solver mode of the oracle (oracle/solver.cpp): parity is UNPINNED by any stored
number of the reference except at t=0; these tests tie it to the pinned mini-app mode and
check its internal consistency across block layouts."""
import numpy as np
import pytest


def test_initial_statistics_match_golden(oracle, golden):
    # gridgeom + qswap + gradcal + statcal of the MAIN solver path at step 0 vs golden row 0
    c = oracle.Case(128, 128, 128)
    c.gridgeom(); c.tgvini(); c.rk_stage(1)
    h = c.history()
    c.close()
    assert abs(h[0, 2] - golden[0, 2]) < 2e-13 * golden[0, 2]
    assert abs(h[0, 3] - golden[0, 3]) < 2e-12 * golden[0, 3]


def test_uniform_metrics(oracle):
    n = 24
    c = oracle.Case(n, n, n)
    c.gridgeom()
    jac = c.get("jacob")[5:-5, 5:-5, 5:-5]
    dx = 2 * np.pi / n
    np.testing.assert_allclose(jac, dx ** 3, rtol=1e-12)
    for a in range(3):
        for b in range(3):
            v = c.get(f"dxi{a + 1}{b + 1}")[5:-5, 5:-5, 5:-5]
            np.testing.assert_allclose(v, (1.0 / dx) if a == b else 0.0, rtol=1e-11, atol=1e-11)
    c.close()


@pytest.mark.parametrize("blocks", [(2, 1, 1), (1, 2, 2), (2, 2, 2)])
def test_block_layouts_agree_to_truncation_level(oracle, blocks):
    # interface closures make results layout dependent but only at truncation level
    n = 32
    ref = oracle.Case(n, n, n)
    ref.gridgeom(); ref.tgvini(); ref.run(2)
    mb = oracle.Case(n, n, n, blocks=blocks)
    mb.gridgeom(); mb.tgvini(); mb.run(2)
    h0, h1 = ref.history(), mb.history()
    assert np.abs(h0[:, 2] - h1[:, 2]).max() < 1e-7
    # and block 0 of the multi-block run covers the same nodes as the single block
    info = mb.block_info(0)
    q_ref = ref.get("q2")[5:5 + info["im"] + 1, 5:5 + info["jm"] + 1, 5:5 + info["km"] + 1]
    q_mb = mb.get("q2", 0)[5:-5, 5:-5, 5:-5]
    assert np.abs(q_ref - q_mb).max() < 1e-6
    assert np.abs(q_ref - q_mb).max() > 0.0
    ref.close(); mb.close()


def test_stage_operators_compose_to_rk_stage(oracle):
    n = 20
    a = oracle.Case(n, n, n); b = oracle.Case(n, n, n)
    for c in (a, b):
        c.gridgeom(); c.tgvini()
    a.rk_stage(1)
    b.filterq(); b.zero_qrhs(); b.qswap(); b.gradcal(); b.save_q(); b.rhscal(); b.rk_update(1); b.updatefvar()
    for name in ("q1", "q2", "q5", "prs", "tmp", "qrhs3"):
        np.testing.assert_array_equal(a.get(name), b.get(name))
    a.close(); b.close()


def test_rk4_is_the_classical_scheme(oracle):
    """rkscheme='rk4' (src/mainloop.F90:368-386, :452-476): q_new = (qsave + dt/6 (k1 + 2 k2 + 2 k3 + k4)) / jacob with
    k_i the qrhs of stage i."""
    n = 16
    c = oracle.Case(n, n, n, deltat=2e-3)
    c.set_flags(lfilter=False, diffterm=True)
    c.set_rkscheme(4)
    c.gridgeom(); c.tgvini()
    ks = []
    for rk in (1, 2, 3, 4):
        c.rk_stage(rk)
        ks.append([c.get(f"qrhs{m + 1}")[5:-5, 5:-5, 5:-5].copy() for m in range(5)])
        if rk == 1:
            qsave = [c.get(f"qsave{m + 1}")[5:-5, 5:-5, 5:-5].copy() for m in range(5)]
    jac = c.get("jacob")[5:-5, 5:-5, 5:-5]
    for m in range(5):
        want = (qsave[m] + 2e-3 / 6.0 * (ks[0][m] + 2 * ks[1][m] + 2 * ks[2][m] + ks[3][m])) / jac
        got = c.get(f"q{m + 1}")[5:-5, 5:-5, 5:-5]
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
    c.close()

    # and a full step through oracle_case_run takes the four stages
    a = oracle.Case(n, n, n, deltat=2e-3); b = oracle.Case(n, n, n, deltat=2e-3)
    for x in (a, b):
        x.set_flags(lfilter=False, diffterm=True); x.set_rkscheme(4); x.gridgeom(); x.tgvini()
    a.run(1)
    for rk in (1, 2, 3, 4):
        b.rk_stage(rk)
    np.testing.assert_array_equal(a.get("q5"), b.get("q5"))
    a.close(); b.close()


def test_global_sponge_damps_outside_the_core_only(oracle):
    """spg_def='circl' (src/sponge_layer.F90:321-440): nodes within range_spange of the centre keep q bit for bit,
    the rest relaxes towards the 6-neighbour mean with a coefficient <= dampfac; two blocks agree with one."""
    n = (24, 20, 16)
    cs = []
    for blocks in ((1, 1, 1), (2, 1, 1)):
        c = oracle.Case(*n, blocks=blocks)
        c.gridgeom(); c.tgvini()
        c.set_sponge_circle(centre=(3.0, 3.0, 3.0), range_spange=2.0, dampfac=0.05)
        cs.append(c)
    one, two = cs
    coef = one.sponge_circle_coef()
    assert coef.max() == pytest.approx(0.05) and coef.min() == 0.0
    before = one.get("q2")[5:-5, 5:-5, 5:-5].copy()
    one.spongefilter(); two.spongefilter()
    after = one.get("q2")[5:-5, 5:-5, 5:-5]
    b = one.block_info(0)
    s_, e_ = b["is_ie"][0::2], b["is_ie"][1::2]
    box = tuple(slice(s_[d], e_[d] + 1) for d in range(3))
    core_mask = coef == 0.0
    np.testing.assert_array_equal(after[box][core_mask], before[box][core_mask])
    assert np.abs(after[box][~core_mask] - before[box][~core_mask]).max() > 0.0
    # the two-block run: same coefficients (global maximum), same result
    left = two.get("q2", 0)[5:-5, 5:-5, 5:-5]
    im0 = two.block_info(0)["im"]
    np.testing.assert_allclose(left[1:im0, :, :], after[1:im0, :, :], rtol=0, atol=1e-15)
    one.close(); two.close()


def test_crash_control_restatement(oracle):
    """lcracon (src/mainloop.F90:709-1198): crashcheck flags negative densities, crashfix replaces sick nodes by the
    mean of their admissible neighbours in storage order, databakup alternates two copies and a repeated recovery of
    the same copy expands the critical nodes to their 27-neighbourhoods."""
    c = oracle.Case(16, 16, 16)
    c.gridgeom(); c.tgvini()
    assert c.crashcheck() == 0 and c.crashfix() == 0
    with pytest.raises(RuntimeError):
        c.databakup("recovery")                      # ' !! not backup data avaliable !!'
    c.databakup("backup"); c.run(1); c.databakup("backup"); c.run(1)
    q = c.get("q1"); q[5 + 3, 5 + 3, 5 + 3] = -1.0; q[5 + 4, 5 + 3, 5 + 3] = -2.0
    c.set("q1", q); c.updatefvar()
    clean = c.get("q1")
    assert c.crashcheck() == 2 and c.get("crinod").sum() == 2.0
    assert c.crashfix() == 2
    fixed = c.get("q1")
    # node (3,3,3) is repaired first from its 25 healthy neighbours, node (4,3,3) then from 25 healthy + the repaired one
    nb = clean[5 + 2:5 + 5, 5 + 2:5 + 5, 5 + 2:5 + 5]
    want = (nb.sum() - clean[8, 8, 8] - clean[9, 8, 8]) / 25.0
    assert fixed[8, 8, 8] == pytest.approx(want, rel=1e-14)
    nb2 = fixed[5 + 3:5 + 6, 5 + 2:5 + 5, 5 + 2:5 + 5]
    assert fixed[9, 8, 8] == pytest.approx((nb2.sum() - fixed[9, 8, 8]) / 26.0, rel=1e-14)
    assert c.get("rho")[9, 8, 8] == fixed[9, 8, 8]
    assert c.nstep == 2
    c.databakup("recovery"); assert c.nstep == 0     # dat_a, the older copy
    c.databakup("recovery"); assert c.nstep == 1     # dat_b
    before = c.get("crinod").sum()
    c.databakup("recovery"); assert c.nstep == 0     # dat_a again: recover_counter 2 -> crinod_expansion
    assert c.get("crinod").sum() > before
    c.close()


def _channel_case(oracle, n=(24, 32, 16), explicit=False, blocks=(1, 1, 1)):
    """examples/Channel/datin/input.chl at reduced size (see tests/gpu_common.py)."""
    sys_path_tests()
    from gpu_common import channel_state, channel_x
    from astr_b200 import refcal
    lengths = (2 * np.pi, 2.0, np.pi)
    c = oracle.Case(*n, blocks=blocks, homo=(True, False, True), reynolds=3000.0, mach=0.3, lengths=lengths)
    c.set_bc((1, 1, 41, 41, 1, 1), (0, 0, 1.0, 1.0, 0, 0))
    c.set_scheme(explicit)
    xg = channel_x(n, lengths)
    for ib in range(c.nblocks):
        b = c.block_info(ib)
        g0, dims = b["g0"], (b["im"], b["jm"], b["km"])
        c.set_x(xg[tuple(slice(g0[d], g0[d] + dims[d] + 1) for d in range(3))], ib)
    c.gridgeom()
    for ib in range(c.nblocks):
        channel_state(c, refcal(3000.0, 0.3), ib=ib)
    return c


def sys_path_tests():
    import os, sys
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)


def test_noslip_wall_closure(oracle):
    # bc.F90:6306: u=0, T=tw, p=(4p1-p2)/3, rho=p/T*const2, q5=p*const6 on the wall nodes only
    from astr_b200 import refcal
    th = refcal(3000.0, 0.3)
    c = _channel_case(oracle)
    before = {k: c.get(k) for k in ("prs", "q1", "u", "tmp")}
    c.boucon()
    for j, j1, j2 in ((5, 6, 7), (-6, -7, -8)):
        p = c.get("prs")
        pe = (4.0 * before["prs"][5:-5, j1, 5:-5] - before["prs"][5:-5, j2, 5:-5]) / 3.0
        np.testing.assert_allclose(p[5:-5, j, 5:-5], pe, rtol=1e-15)
        assert np.all(c.get("u")[5:-5, j, 5:-5] == 0.0) and np.all(c.get("q3")[5:-5, j, 5:-5] == 0.0)
        assert np.all(c.get("tmp")[5:-5, j, 5:-5] == 1.0)
        np.testing.assert_allclose(c.get("q1")[5:-5, j, 5:-5], pe * th["const2"], rtol=1e-15)
        np.testing.assert_allclose(c.get("q5")[5:-5, j, 5:-5], pe * th["const6"], rtol=1e-15)
    # interior untouched
    np.testing.assert_array_equal(c.get("q1")[5:-5, 6:-6, 5:-5], before["q1"][5:-5, 6:-6, 5:-5])
    c.close()


def test_src_chan_adds_force_times_jacobian(oracle):
    # solver.F90:341-352: qrhs(2:4)+=force*J, qrhs(5)+=force.ubulk*J ; Poiseuille: ubulk ~ 1
    a, b = _channel_case(oracle), _channel_case(oracle)
    force = (2.5e-3, 0.0, 1e-4)
    a.set_flow(1, force); b.set_flow(0, force)
    for c in (a, b):
        c.boucon(); c.qswap(); c.gradcal(); c.zero_qrhs(); c.rhscal()
    jac = a.get("jacob")[5:-5, 5:-5, 5:-5]
    d2 = (a.get("qrhs2") - b.get("qrhs2"))[5:-5, 5:-5, 5:-5]
    d4 = (a.get("qrhs4") - b.get("qrhs4"))[5:-5, 5:-5, 5:-5]
    d5 = (a.get("qrhs5") - b.get("qrhs5"))[5:-5, 5:-5, 5:-5]
    scale = np.abs(a.get("qrhs2")).max()
    assert np.abs(d2 - force[0] * jac).max() < 1e-14 * scale
    assert np.abs(d4 - force[2] * jac).max() < 1e-14 * scale
    ub = d5 / jac
    assert np.ptp(ub) < 1e-10 and abs(ub.mean() / force[0] - 1.0) < 0.05   # u1bulk of 1.5(1-eta^2) is 1
    np.testing.assert_array_equal(a.get("qrhs1"), b.get("qrhs1"))
    a.close(); b.close()


@pytest.mark.parametrize("explicit", [False, True])
def test_channel_runs_and_split_in_y_agrees(oracle, explicit):
    # wall closures (ntype 4 single block; 1/2 when split in y) + noslip + src_chan stay stable and the
    # two-block layout differs only at truncation level (psum of the bulk integrals over blocks)
    one, two = _channel_case(oracle, explicit=explicit), _channel_case(oracle, explicit=explicit, blocks=(1, 2, 1))
    for c in (one, two):
        c.set_flow(1, (2.5e-3, 0.0, 0.0))
        for rk in (1, 2, 3):
            c.rk_stage(rk)
    jm0 = two.block_info(0)["jm"]
    u1 = one.get("u")[5:-5, 5:5 + jm0 + 1, 5:-5]
    u2 = two.get("u", 0)[5:-5, 5:-5, 5:-5]
    assert np.isfinite(u1).all() and np.abs(u1 - u2).max() < 1e-4
    one.close(); two.close()


def test_explicit_scheme_is_sixth_order_on_the_rhs(oracle):
    # diff6ec vs df_compact on a smooth periodic field: both 6th order -> qrhs agree to truncation level
    n = 32
    a, b = oracle.Case(n, n, n), oracle.Case(n, n, n)
    b.set_scheme(True)
    for c in (a, b):
        c.gridgeom(); c.tgvini(); c.qswap(); c.gradcal(); c.zero_qrhs(); c.rhscal()
    ra, rb = a.get("qrhs2")[5:-5, 5:-5, 5:-5], b.get("qrhs2")[5:-5, 5:-5, 5:-5]
    err = np.abs(ra - rb).max() / np.abs(ra).max()
    assert 0.0 < err < 1e-4
    a.close(); b.close()


def test_open_face_boundary_conditions(oracle):
    # inflow(1) bc.F90:1366, outflow(2) :3404, farfield(4) :3008 on a block with walls in i and j
    sys_path_tests()
    from gpu_common import stretched_x
    n, homo = (24, 20, 12), (False, False, True)
    c = oracle.Case(*n, homo=homo)
    c.set_bc((11, 21, 41, 51, 1, 1), (0, 0, 1.05, 0, 0, 0))
    c.set_x(stretched_x(n, homo)); c.gridgeom(); c.tgvini()
    jm, km = n[1], n[2]
    vel_in = np.zeros((jm + 1, km + 1, 3), order="F")
    vel_in[:, :, 0] = 40.0                     # Mach 4 against css = sqrt(T)/M = 10: blend -> 1
    tmp_in = np.ones((jm + 1, km + 1), order="F")
    c.set_inflow(vel_in, tmp_in, np.ones(jm + 1))
    # a linear-in-j profile: extrapolate(v1,v2,dv=0) is the one-sided zero-gradient closure (4 v1 - v2)/3
    for name, slope in (("u", 0.01), ("prs", 0.2), ("rho", 0.003)):
        a = c.get(name)
        a[...] = (2.0 if name != "prs" else 70.0) + slope * np.arange(a.shape[1])[None, :, None]
        c.set(name, a)
    before = {k: c.get(k) for k in ("u", "prs", "rho", "tmp")}
    c.boucon()
    u, p, r = c.get("u"), c.get("prs"), c.get("rho")
    # farfield at jmax (away from the i faces, which outflow / inflow own afterwards... boucon order is n=1..6)
    jt = 5 + jm
    for name, a in (("u", u), ("prs", p), ("rho", r)):
        want = (4.0 * before[name][8:-8, jt - 1, 5:-5] - before[name][8:-8, jt - 2, 5:-5]) / 3.0
        np.testing.assert_allclose(a[8:-8, jt, 5:-5], want, rtol=1e-14)
    # supersonic inflow: p = pinf, u = vel_in
    np.testing.assert_allclose(p[5, 6:-6, 5:-5], c.pinf, rtol=1e-9)
    np.testing.assert_allclose(u[5, 6:-6, 5:-5], 40.0, rtol=1e-9)
    # outflow at imax: first-order copy of the neighbour plane
    np.testing.assert_array_equal(p[-6, 6:-6, 5:-5], p[-7, 6:-6, 5:-5])
    # conserved variables consistent with the primitives on the treated faces
    q1 = c.get("q1")
    np.testing.assert_array_equal(q1[5, 6:-6, 5:-5], r[5, 6:-6, 5:-5])
    np.testing.assert_array_equal(q1[8:-8, jt, 5:-5], r[8:-8, jt, 5:-5])
    c.close()
