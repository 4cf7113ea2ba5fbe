"""Solution-level checks of the tracer path, in the spirit of the reference's own convergence tests (SURVEY.md
section 4: the reference pins this path only at solution level). The assembled system is used exactly as
solve_field_equation_cg uses it (assemble/Advection_Diffusion_CG.F90:127-175): matrix . delta_T = rhs for the RATE
of change, strong Dirichlet rows carry (value - T)/dt and are lifted out of the solve
(femtools/Boundary_Conditions.F90:1982-2024, femtools/Solvers.F90:1072-1098), then T += dt * delta_T (:1396-1405).
Analytic solutions of the heat and advection-diffusion equations must be approached at second order in h
(Crank-Nicolson keeps the time error below the space error here). CPU only (oracle + scipy)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from fluidity_b200 import synthetic as syn, _abi as abi


def advance(orc, mesh, fs, opts, findrm, colm, T, dt, dirichlet=None):
    fs.set(abi.F_T, T)
    sys_ = orc.assemble_advdiff(mesh, fs, opts, findrm, colm)
    n = mesh.n_nodes
    A = sp.csr_matrix((sys_["matrix"], colm - 1, findrm - 1), shape=(n, n))
    rhs = sys_["rhs"]
    delta = np.zeros(n)
    active = np.ones(n, dtype=bool)
    if dirichlet is not None:
        nodes, values = dirichlet
        inactive = np.zeros(n, dtype=np.int32)
        orc.apply_dirichlet_scalar(nodes, values, T, dt, rhs, inactive)
        active = inactive == 0
        delta[~active] = rhs[~active]                       # "the right value will be substituted after the solve"
        rhs = rhs - A @ np.where(active, 0.0, delta)        # lifting of the ghost columns
    a = np.flatnonzero(active)
    delta[a] = spla.spsolve(A[a][:, a].tocsc(), rhs[a])
    return T + dt * delta


def l2_error(mesh, T, exact):
    # lumped-mass L2 norm
    Xe = mesh.X[mesh.ndglno - 1]
    vol = np.abs(np.linalg.det(Xe[:, 1:] - Xe[:, :1])) / (2 if mesh.dim == 2 else 6)
    w = np.zeros(mesh.n_nodes)
    np.add.at(w, (mesh.ndglno - 1).ravel(), np.repeat(vol / mesh.loc, mesh.loc))
    return float(np.sqrt((w * (T - exact) ** 2).sum()))


def test_heat_equation_decay_converges_at_second_order(orc):
    """dT/dt = kappa Lap T on the unit square, T0 = cos(pi x): natural (zero-flux) boundaries, no face terms.
    Exact: cos(pi x) exp(-kappa pi^2 t)."""
    kappa, dt, t_end = 0.5, 0.0025, 0.05
    errs = []
    for n in (8, 16, 32):
        mesh = syn.box_mesh((n, n), seed=11)
        fs = syn.standard_fields(mesh)
        fs.set(abi.F_T_DIFFUSIVITY, syn.iso_tensor(2, kappa), abi.FIELD_CONSTANT)
        findrm, colm, _ = orc.make_sparsity(mesh)
        o = abi.common_advdiff_opts(have_advection=0, dt=dt, theta=0.5)
        T = np.cos(np.pi * mesh.X[:, 0])
        for _ in range(int(round(t_end / dt))):
            T = advance(orc, mesh, fs, o, findrm, colm, T, dt)
        errs.append(l2_error(mesh, T, np.cos(np.pi * mesh.X[:, 0]) * np.exp(-kappa * np.pi ** 2 * t_end)))
    rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert errs[-1] < 2e-4 and (rates > 1.8).all(), (errs, rates)


@pytest.mark.parametrize("lumped", [0, 1])
def test_advection_diffusion_wave_with_strong_dirichlet_converges(orc, lumped):
    """dT/dt + u.grad T = kappa Lap T with constant u = (1, 0): T = exp(-kappa k^2 t) cos(k (x - u t)), imposed on
    the inflow/outflow walls as a strong Dirichlet condition; the y-walls are natural (the solution does not
    depend on y)."""
    kappa, k, dt, t_end = 0.05, 2.0 * np.pi, 0.002, 0.1
    exact = lambda X, t: np.exp(-kappa * k * k * t) * np.cos(k * (X[:, 0] - t))
    errs = []
    for n in (8, 16, 32):
        mesh = syn.box_mesh((n, n), seed=12)
        fs = syn.standard_fields(mesh)
        u = np.zeros((mesh.n_nodes, 2))
        u[:, 0] = 1.0
        fs.set(abi.F_NU, u)
        fs.set(abi.F_T_DIFFUSIVITY, syn.iso_tensor(2, kappa), abi.FIELD_CONSTANT)
        findrm, colm, _ = orc.make_sparsity(mesh)
        o = abi.common_advdiff_opts(dt=dt, theta=0.5, lump_mass=lumped)
        walls = np.flatnonzero((mesh.X[:, 0] < 1e-12) | (mesh.X[:, 0] > 1 - 1e-12)) + 1
        T = exact(mesh.X, 0.0)
        t = 0.0
        for _ in range(int(round(t_end / dt))):
            T = advance(orc, mesh, fs, o, findrm, colm, T, dt, (walls, exact(mesh.X[walls - 1], t + dt)))
            t += dt
        assert np.abs(T[walls - 1] - exact(mesh.X[walls - 1], t)).max() < 1e-12   # Dirichlet rows hold the value
        errs.append(l2_error(mesh, T, exact(mesh.X, t)))
    rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert errs[-1] < 5e-3 and (rates > 1.7).all(), (errs, rates)


@pytest.mark.parametrize("lump", [1, 0])
def test_momentum_shear_layer_decay_converges_at_second_order(orc, lump):
    """Momentum equation without pressure gradient: u = (cos(pi y) exp(-nu pi^2 t), 0) is an exact solution
    (u.grad u = 0, zero-stress walls are natural). big_m . delta_u = rhs per component for the rate of change
    (assemble/Momentum_Equation.F90: the velocity change is solved for and added with dt), advecting velocity and
    oldu = u^n, constant density."""
    nu_, dt, t_end = 0.2, 0.0025, 0.05
    errs = []
    for n in (8, 16, 32):
        mesh = syn.box_mesh((n, n), seed=13)
        fs = syn.standard_fields(mesh)
        fs.set(abi.F_DENSITY, np.array([1.0]), abi.FIELD_CONSTANT)
        fs.set(abi.F_VISCOSITY, syn.iso_tensor(2, nu_), abi.FIELD_CONSTANT)
        findrm, colm, _ = orc.make_sparsity(mesh)
        o = abi.common_momentum_opts(dt=dt, theta=0.5, have_gravity=0, lump_mass=lump)
        N = mesh.n_nodes
        u = np.zeros((N, 2))
        u[:, 0] = np.cos(np.pi * mesh.X[:, 1])
        for _ in range(int(round(t_end / dt))):
            fs.set(abi.F_NU, u)
            fs.set(abi.F_OLDU, u)
            s = orc.assemble_momentum(mesh, fs, o, findrm, colm)
            for d in range(2):
                A = sp.csr_matrix((s["big_m"][d], colm - 1, findrm - 1), shape=(N, N)).tocsc()
                u[:, d] = u[:, d] + dt * spla.spsolve(A, s["rhs"][:, d])
        exact = np.cos(np.pi * mesh.X[:, 1]) * np.exp(-nu_ * np.pi ** 2 * t_end)
        errs.append(l2_error(mesh, u[:, 0], exact))
        assert np.abs(u[:, 1]).max() < 1e-10       # no spurious cross-flow
    rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert errs[-1] < 3e-4 and (rates > 1.8).all(), (errs, rates)


def test_pressure_matrix_is_a_consistent_discrete_laplacian(orc):
    """ct_m(d)(i, j) = int N_i d_d N_j (Momentum_CG.F90:1401) is the weak divergence; for p vanishing on the
    boundary C^T p = -int grad p_h N_j, so cmc_m = C M_L^-1 C^T (Assemble_CMC.F90:119-135) is the stiffness form
    q^T cmc p -> int grad q . grad p, at second order on jittered meshes (pointwise it is only first order there:
    the stencil is two elements wide). p = sin(pi x) sin(2 pi y), q = x(1-x) sin(2 pi y):
    int |grad p|^2 = 5 pi^2 / 4, int grad q . grad p = 10 / pi."""
    errs = []
    for n in (8, 16, 32):
        mesh = syn.box_mesh((n, n), seed=14)
        fs = syn.standard_fields(mesh)
        fs.set(abi.F_DENSITY, np.array([1.0]), abi.FIELD_CONSTANT)
        findrm, colm, _ = orc.make_sparsity(mesh)
        s = orc.assemble_momentum(mesh, fs, abi.common_momentum_opts(assemble_ct_matrix_here=1), findrm, colm, want_ct=True)
        f2, c2 = orc.make_sparsity_mult(mesh.n_nodes, findrm, colm)
        cmc = orc.mult_div_vector_div_T(findrm, colm, s["ct_m"], s["ct_m"], 1.0 / s["masslump"], f2, c2)
        N = mesh.n_nodes
        M = sp.csr_matrix((cmc, c2 - 1, f2 - 1), shape=(N, N))
        x, y = mesh.X[:, 0], mesh.X[:, 1]
        p = np.sin(np.pi * x) * np.sin(2 * np.pi * y)
        q = x * (1 - x) * np.sin(2 * np.pi * y)
        errs.append(max(abs(p @ (M @ p) - 5 * np.pi ** 2 / 4) / (5 * np.pi ** 2 / 4), abs(q @ (M @ p) - 10 / np.pi) / (10 / np.pi)))
    rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert errs[-1] < 0.02 and (rates > 1.8).all(), (errs, rates)
