"""Option variants shared by the parity tests (CPU oracle + CUDA path) and by the golden-vector generator
(tests/golden/make_pyref_golden.py): one table, so every non-stabilised variant the GPU tests run has a `c_*` golden
computed with the reference's own Python element loops."""
import numpy as np

from fluidity_b200 import synthetic as syn, _abi as abi


def momentum_variants():
    c = abi.common_momentum_opts
    return {
        "common": c(),
        "common_ct": c(assemble_ct_matrix_here=1),
        "consistent_mass": c(lump_mass=0),
        "by_parts_beta": c(integrate_advection_by_parts=1, beta=0.3),
        "beta1": c(beta=1.0),
        "absorption": c(have_absorption=1),
        "absorption_lumped_pc": c(have_absorption=1, lump_absorption=1, pressure_corrected_absorption=1),
        "source": c(have_source=1),
        "source_lumped": c(have_source=1, lump_source=1),
        "ref_profile": c(subtract_out_reference_profile=1),
        "aniso": c(viscosity_shape=abi.TENSOR_FULL),
        "diagvisc": c(viscosity_shape=abi.TENSOR_DIAGONAL),
        "no_adv_no_mass": c(exclude_advection=1, exclude_mass=1),
        "stokes_no_ml": c(exclude_advection=1, have_gravity=0, assemble_inverse_masslump=0),
        "su_optimal": c(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND, nu_bar_scheme=abi.NU_BAR_OPTIMAL),
        "su_unity_noviscosity": c(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND, have_viscosity=0),
        "supg_critical": c(stabilisation_scheme=abi.STAB_SUPG, nu_bar_scheme=abi.NU_BAR_CRITICAL_RULE, nu_bar_scale=1.0,
                           lump_mass=0, have_absorption=1, have_source=1),
        "supg_asymptotic_by_parts": c(stabilisation_scheme=abi.STAB_SUPG, nu_bar_scheme=abi.NU_BAR_DOUBLY_ASYMPTOTIC,
                                      integrate_advection_by_parts=1, beta=0.5),
    }


def advdiff_variants():
    c = abi.common_advdiff_opts
    return {
        "common": c(),
        "lumped": c(lump_mass=1),
        "by_parts": c(integrate_advection_by_parts=1, beta=0.25),
        "beta": c(beta=1.0),
        "absorb_source": c(have_absorption=1, have_source=1),
        "tensor_diff": c(diffusivity_shape=abi.TENSOR_FULL),
        "pure_diffusion": c(have_advection=0),
        "mass_only": c(have_advection=0, have_diffusivity=0),
        "theta0": c(theta=0.0),
        "su_optimal": c(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND),
        "su_unity_nodiff": c(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND, have_diffusivity=0),
        "supg_optimal_tensor": c(stabilisation_scheme=abi.STAB_SUPG, diffusivity_shape=abi.TENSOR_FULL, have_source=1,
                                 have_absorption=1, lump_mass=1),
        "supg_critical_by_parts": c(stabilisation_scheme=abi.STAB_SUPG, nu_bar_scheme=abi.NU_BAR_CRITICAL_RULE,
                                    integrate_advection_by_parts=1, beta=0.3),
    }


def fields_for(mesh, variant):
    fs = syn.standard_fields(mesh, nodal_viscosity=(variant == "aniso"))
    if variant in ("diagvisc", "tensor_diff", "supg_optimal_tensor"):
        fs.set(abi.F_VISCOSITY, syn.aniso_tensor(mesh.dim), abi.FIELD_CONSTANT)
        fs.set(abi.F_T_DIFFUSIVITY, syn.aniso_tensor(mesh.dim), abi.FIELD_CONSTANT)
    return fs



def boussinesq_variants():
    c = abi.common_momentum_opts
    return {
        "absorption": c(have_absorption=1),
        "absorption_nogravity": c(have_absorption=1, have_gravity=0),          # backward_facing_step_3d's option set
        "absorption_lumped": c(have_absorption=1, lump_absorption=1),
        "absorption_lumped_pc": c(have_absorption=1, lump_absorption=1, pressure_corrected_absorption=1),
        "absorption_pc_full": c(have_absorption=1, pressure_corrected_absorption=1),
        "source": c(have_source=1),
        "source_lumped": c(have_source=1, lump_source=1),
        "ref_profile": c(subtract_out_reference_profile=1),
        "everything": c(have_absorption=1, have_source=1, subtract_out_reference_profile=1, viscosity_shape=abi.TENSOR_FULL),
        "everything_lumped_noml": c(have_absorption=1, lump_absorption=1, pressure_corrected_absorption=1, have_source=1,
                                    lump_source=1, subtract_out_reference_profile=1, assemble_inverse_masslump=0),
        "exclude_mass_adv": c(have_absorption=1, have_source=1, exclude_mass=1, exclude_advection=1),
    }


# ---- the variants that have a golden from the reference's own Python loops (tests/golden/make_pyref_golden.py) ------
def variant_cases():
    """(tag, kind, opts, field tag) of every NON-STABILISED variant above: what has `v_*` golden vectors. SU / SUPG have
    no reference implementation outside the Fortran and stay pinned by the oracle's closed forms only."""
    out = []
    for tag, o in momentum_variants().items():
        if not o.stabilisation_scheme:
            out.append((tag, "mom", o, tag))
    for tag, o in boussinesq_variants().items():
        out.append(("bq_" + tag, "mom", o, "bq_" + tag))
    for tag, o in advdiff_variants().items():
        if not o.stabilisation_scheme:
            out.append((tag, "adv", o, tag))
    return out


def variant_field_key(ftag):
    """Variants with the same key share one field set."""
    if ftag == "aniso":
        return "aniso"
    if ftag in ("diagvisc", "tensor_diff"):
        return "tensor"
    if ftag == "bq_everything":
        return "bq_everything"
    return "bq" if ftag.startswith("bq_") else "std"


def variant_fields(mesh, ftag):
    """The field set the parity tests use for the variant: fields_for, and constant density 1.3 for the Boussinesq sets
    (test_strip_additive_pass_constant_density)."""
    if ftag.startswith("bq_"):
        fs = syn.standard_fields(mesh)
        fs.set(abi.F_DENSITY, np.array([1.3]), abi.FIELD_CONSTANT)
        if ftag == "bq_everything":
            fs.set(abi.F_VISCOSITY, syn.aniso_tensor(mesh.dim), abi.FIELD_CONSTANT)
        return fs
    return fields_for(mesh, ftag)
