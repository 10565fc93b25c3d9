"""Device post-processing (csrc/postproc.cu) against the numpy restatements in oracle/postproc.py:
FreeSurface<3>::compute_internal_velocities (free_surface.cc:10426-10537) and the hull integrals of
compute_pressure (free_surface.cc:9534-9598).  Tolerance 1e-8 relative (north_star: hull drag /
potential 1e-8); in practice ~1e-13."""
import numpy as np
import pytest

from oracle import postproc
from wavebem_b200 import meshgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solved(wb):
    m = meshgen.wigley_tank(nxm=12, nt=6, nxu=4, nxd=6, nz=3, nzh=3, wave_amp=0.02, wave_phase=0.4)
    bc = meshgen.towing_tank_bc(m)
    ctx = wb.Context(gmres_tol=1e-12, gmres_max_steps=400, auto_constraints=1)   # compute_constraints inside the library
    ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    s = m.surface_nodes == 1
    z = np.zeros(m.n_nodes)
    phi, dphi, _, _ = ctx.solve(m.xyz, np.where(s, bc, 0.0), np.where(s, 0.0, bc), bc)
    yield dict(m=m, ctx=ctx, phi=phi, dphi=dphi)
    ctx.close()


def _field_points(m, n=37):
    lo, hi = m.xyz.min(axis=0), m.xyz.max(axis=0)
    rng = np.random.default_rng(11)
    # inside the tank, below the free surface, away from the boundary
    return lo + (0.2 + 0.6 * rng.random((n, 3))) * (hi - lo) * np.array([1.0, 1.0, 0.5])


def test_internal_velocities_match_the_restatement(solved):
    m, ctx = solved["m"], solved["ctx"]
    pts = _field_points(m)
    ref = postproc.internal_velocities(m, solved["phi"], solved["dphi"], pts)
    got = ctx.internal_velocities(solved["phi"], solved["dphi"], pts)
    assert np.abs(got - ref).max() <= 1e-8 * np.abs(ref).max()
    # many points (several CTAs, a ragged last one) and a single point
    pts2 = _field_points(m, 300)
    got2 = ctx.internal_velocities(solved["phi"], solved["dphi"], pts2)
    assert np.array_equal(got2[:5], ctx.internal_velocities(solved["phi"], solved["dphi"], pts2[:5]))
    ref2 = postproc.internal_velocities(m, solved["phi"], solved["dphi"], pts2[::29])
    assert np.abs(got2[::29] - ref2).max() <= 1e-8 * np.abs(ref2).max()
    assert ctx.internal_velocities(solved["phi"], solved["dphi"], np.zeros((0, 3))).shape == (0, 3)


def test_internal_velocity_of_a_linear_potential(wb):
    """Green's representation of phi = x inside a closed cube: grad phi = (1, 0, 0)."""
    m = meshgen.cube(8)
    nn = meshgen.cell_normals_at_nodes(m)
    ctx = wb.Context()
    ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    ctx.set_geometry(m.xyz)
    pts = np.array([[0.0, 0.0, 0.0], [0.1, -0.15, 0.05]]) + m.xyz.mean(axis=0)
    v = ctx.internal_velocities(m.xyz[:, 0], nn[:, 0], pts)
    assert np.abs(v - np.array([1.0, 0.0, 0.0])).max() < 2e-3
    ctx.close()


def test_pressure_force_matches_the_restatement(solved):
    m, ctx = solved["m"], solved["ctx"]
    vinf = np.array([0.28 * np.sqrt(9.81 * meshgen.WIGLEY_L), 0.0, 0.0])
    marked = np.isin(m.cell_patch, m.meta["hull_patches"]).astype(np.uint8)
    bar = (0.3, 0.0, -0.05)
    ref = postproc.hull_pressure_force(m, solved["phi"], solved["dphi"], vinf, full=True, baricenter=bar)
    got = ctx.pressure_force(solved["phi"], solved["dphi"], marked, vinf, baricenter=bar)
    scale = np.array([np.abs(ref[0:3]).max()] * 3 + [np.abs(ref[3:6]).max()] * 3 + [np.abs(ref[6:9]).max()] * 3 +
                     [ref[9], abs(ref[10]) + ref[9]])
    assert (np.abs(got - ref) / scale).max() < 1e-8
    # nothing marked -> zeros; everything marked -> the closed-surface area
    assert np.array_equal(ctx.pressure_force(solved["phi"], solved["dphi"], np.zeros(m.n_cells, np.uint8), vinf), np.zeros(11))
    allm = ctx.pressure_force(solved["phi"], solved["dphi"], np.ones(m.n_cells, np.uint8), vinf)
    ref_all = postproc.hull_pressure_force(m, solved["phi"], solved["dphi"], vinf, full=True,
                                           hull_patches=np.unique(m.cell_patch))
    assert abs(allm[9] - ref_all[9]) < 1e-10 * ref_all[9]


def test_postprocessing_through_a_single_process_group(wb, solved):
    m = solved["m"]
    g = wb.Context(n_gpus=2, devices=[0, 0])
    g.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    g.set_geometry(m.xyz)
    pts = _field_points(m, 9)
    assert np.array_equal(g.internal_velocities(solved["phi"], solved["dphi"], pts),
                          solved["ctx"].internal_velocities(solved["phi"], solved["dphi"], pts))
    g.close()
