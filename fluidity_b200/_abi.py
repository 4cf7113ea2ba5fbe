"""ctypes mirror of include/cgasm.h (enums + option structs).

Field order and types must match the header exactly; tests/test_abi.py checks sizeof and
the exported symbol list against the header text.
"""
import ctypes as C

# status codes
OK, EHANDLE, EARG, EUNSUPPORTED, ESTATE, ECUDA, ENCCL, ENODEVICE = range(8)

# field slots
(F_NU, F_OLDU, F_DENSITY, F_VISCOSITY, F_BUOYANCY, F_HB_DENSITY, F_GRAVITY, F_ABSORPTION,
 F_SOURCE, F_T, F_T_DIFFUSIVITY, F_T_SOURCE, F_T_ABSORPTION, F_NSLOTS) = range(14)

FIELD_RANK = {
    F_NU: 1, F_OLDU: 1, F_DENSITY: 0, F_VISCOSITY: 2, F_BUOYANCY: 0, F_HB_DENSITY: 0,
    F_GRAVITY: 1, F_ABSORPTION: 1, F_SOURCE: 1, F_T: 0, F_T_DIFFUSIVITY: 2, F_T_SOURCE: 0,
    F_T_ABSORPTION: 0,
}

FIELD_NORMAL, FIELD_CONSTANT = 0, 1
STAB_NONE, STAB_STREAMLINE_UPWIND, STAB_SUPG = 0, 1, 2
NU_BAR_OPTIMAL, NU_BAR_DOUBLY_ASYMPTOTIC, NU_BAR_CRITICAL_RULE, NU_BAR_UNITY = 1, 2, 3, 4
TENSOR_ISOTROPIC, TENSOR_DIAGONAL, TENSOR_FULL = 0, 1, 2
SCATTER_ATOMIC, SCATTER_COLOURED, SCATTER_WARPAGG, SCATTER_TILED, SCATTER_GATHER, SCATTER_STRIP = 0, 1, 2, 3, 4, 5
# boundary-condition types of the surface loops (Advection_Diffusion_CG.F90:74-75, Momentum_CG.F90:138-140)
TBC_NONE, TBC_NEUMANN, TBC_WEAKDIRICHLET, TBC_INTERNAL, TBC_ROBIN = range(5)
VBC_NONE, VBC_WEAKDIRICHLET, VBC_NO_NORMAL_FLOW, VBC_INTERNAL, VBC_FREE_SURFACE, VBC_FLUX = range(6)

_M_DOUBLES = ["dt", "theta", "beta", "gravity_magnitude", "nu_bar_scale", "fs_sf"]
_M_INTS = [
    "lump_mass", "exclude_mass", "exclude_advection", "integrate_advection_by_parts",
    "have_source", "lump_source", "have_gravity", "subtract_out_reference_profile",
    "have_absorption", "lump_absorption", "pressure_corrected_absorption", "have_viscosity",
    "viscosity_shape", "assemble_inverse_masslump", "assemble_ct_matrix_here",
    "stabilisation_scheme", "nu_bar_scheme",
    # unsupported switches
    "have_les", "multiphase", "on_sphere", "move_mesh", "have_coriolis",
    "have_geostrophic_pressure", "have_surfacetension", "have_vertical_stabilization",
    "have_swe_bottom_drag", "have_wd_abs", "have_temperature_dependent_viscosity",
    "stress_form", "partial_stress_form", "radial_gravity", "vel_lump_on_submesh",
    "cmc_lump_on_submesh", "abs_lump_on_submesh",
    # implemented since round 2
    "assemble_mass_matrix", "integrate_continuity_by_parts",
    # surface loop only: free-surface stabilisation (scale factor fs_sf)
    "have_surface_fs_stabilisation",
]


class MomentumOpts(C.Structure):
    """cgasm_momentum_opts: module-level switches of assemble/Momentum_CG.F90:83-178."""
    _fields_ = [(k, C.c_double) for k in _M_DOUBLES] + [(k, C.c_int) for k in _M_INTS]

    def __init__(self, **kw):
        super().__init__()
        self.nu_bar_scheme = NU_BAR_OPTIMAL
        self.nu_bar_scale = 0.5
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


_A_DOUBLES = ["dt", "theta", "beta", "nu_bar_scale"]
_A_INTS = [
    "have_mass", "lump_mass", "have_advection", "integrate_advection_by_parts", "have_source",
    "have_absorption", "have_diffusivity", "diffusivity_shape", "stabilisation_scheme",
    "nu_bar_scheme", "move_mesh", "multiphase", "equation_type_not_advdiff",
]


class AdvDiffOpts(C.Structure):
    """cgasm_advdiff_opts: switches of assemble/Advection_Diffusion_CG.F90:77-123."""
    _fields_ = [(k, C.c_double) for k in _A_DOUBLES] + [(k, C.c_int) for k in _A_INTS]

    def __init__(self, **kw):
        super().__init__()
        self.nu_bar_scheme = NU_BAR_OPTIMAL
        self.nu_bar_scale = 0.5
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


def common_momentum_opts(**kw):
    """The option set shared by the four example configs:
    lump_mass, tensor-form isotropic viscosity, no stabilisation, beta=0, advection not by
    parts, theta=0.5; inverse lumped mass assembled (Momentum_CG.F90:847-878)."""
    base = dict(dt=0.01, theta=0.5, beta=0.0, gravity_magnitude=10.0, lump_mass=1,
                have_gravity=1, have_viscosity=1, viscosity_shape=TENSOR_ISOTROPIC,
                assemble_inverse_masslump=1)
    base.update(kw)
    return MomentumOpts(**base)


def common_advdiff_opts(**kw):
    """Default CG tracer: consistent mass, advection not by parts, isotropic diffusivity."""
    base = dict(dt=0.01, theta=0.5, beta=0.0, have_mass=1, have_advection=1,
                have_diffusivity=1, diffusivity_shape=TENSOR_ISOTROPIC)
    base.update(kw)
    return AdvDiffOpts(**base)
