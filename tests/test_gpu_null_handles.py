"""Every species / grid entry point of the C ABI that dereferences its handle answers a NULL handle with PGPU_ERR_ARG
(or PGPU_ERR_STATE) instead of crashing: the reference-side shim can then report the error the way MayDay::Error would."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu

SPECIES_CALLS = {
    "pgpu_advance_positions_explicit": (C.c_double(0.1), C.c_int(0)),
    "pgpu_advance_positions_implicit": (C.c_double(0.1),),
    "pgpu_advance_positions_2nd_half": (),
    "pgpu_advance_velocities_2nd_half": (),
    "pgpu_average_velocities": (),
    "pgpu_update_old_particle_positions": (),
    "pgpu_update_old_particle_velocities": (),
    "pgpu_interpolate_fields_to_particles": (),
    "pgpu_set_moments_from_bins": (),
    "pgpu_bin_particles": (),
    "pgpu_sort_for_locality": (),
}


@pytest.mark.parametrize("name", sorted(SPECIES_CALLS))
def test_null_species_handle_is_an_argument_error(pgpu, name):
    fn = getattr(pgpu.load(), name)
    fn.restype = C.c_int
    rc = fn(C.c_void_p(None), *SPECIES_CALLS[name])
    assert rc in (-1, -5), (name, rc)          # PGPU_ERR_ARG / PGPU_ERR_STATE


def test_null_handles_in_the_collision_and_density_calls(pgpu):
    lib = pgpu.load()
    null = C.c_void_p(None)
    prm = pgpu.CoulombParams(10.0, 1, 0, 11, 1)
    assert lib.pgpu_collide_coulomb(null, null, C.byref(prm), C.c_double(1e-12), C.c_uint64(1), C.c_uint64(0), None) < 0
    assert lib.pgpu_collide_ta(null, null, C.c_double(3.0), C.c_double(1e-12), C.c_uint64(1), C.c_uint64(0), None) < 0
    assert lib.pgpu_collide_hard_sphere(null, null, C.c_double(1e-19), C.c_double(1e-12), C.c_uint64(1), C.c_uint64(0),
                                        None) < 0
    st = (C.c_int * 2)(1, 1)
    assert lib.pgpu_charge_density_deposit(null, st) < 0
    assert lib.pgpu_charge_density_filter(null, st) < 0
    assert lib.pgpu_current_filter(null, 1, 1) < 0
    out = C.c_void_p()
    assert lib.pgpu_halo_create_rho(null, st, 0, None, C.byref(out)) < 0
