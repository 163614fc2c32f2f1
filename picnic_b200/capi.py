"""ctypes binding of the C ABI (include/picnic_gpu.h) -- used by tests/ and bench.py.

This is the same boundary the C++ shim (picnic_b200/host/) calls.  There is no CPU
fallback: if libpicnic_gpu.so is missing, or no CUDA device is usable, every entry
point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpicnic_gpu.so")

CIC, TSC, CC0, CC1 = 0, 1, 2, 3
BC_NONE, BC_PERIODIC, BC_SYMMETRY = 0, 1, 2
ERR_SEGMENTS, ERR_BOUNDS = -3, -4


class PgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pgpu error %d: %s" % (code, msg))
        self.code = code


class ExtFn(C.Structure):
    """pgpu_ext_fn (include/picnic_gpu.h): one external-field grid function."""
    _fields_ = [("type", C.c_int), ("value", C.c_double), ("constant", C.c_double),
                ("L", C.c_double * 2), ("mode", C.c_double * 2), ("phase", C.c_double * 2),
                ("C", C.c_double * 2), ("A", C.c_double * 2), ("X0", C.c_double * 2), ("eps", C.c_double * 2)]


class GridDesc(C.Structure):
    _fields_ = [("D", C.c_int), ("ncell", C.c_int * 2), ("xmin", C.c_double * 2), ("dx", C.c_double * 2),
                ("nghost", C.c_int), ("periodic", C.c_int * 2), ("box_lo", C.c_int * 2),
                ("box_hi", C.c_int * 2), ("volume_scale", C.c_double)]


class SpeciesDesc(C.Structure):
    _fields_ = [("mass", C.c_double), ("charge", C.c_double), ("fnorm_const", C.c_double),
                ("cvac_norm", C.c_double), ("interp_N", C.c_int), ("interp_J", C.c_int),
                ("interp_E", C.c_int), ("rtol", C.c_double), ("iter_max", C.c_int),
                ("order_swap", C.c_int), ("bc_check_lo", C.c_int * 2), ("bc_check_hi", C.c_int * 2),
                ("motion", C.c_int), ("forces", C.c_int), ("relativistic", C.c_int), ("higuera_cary", C.c_int)]


# particle boundary conditions of pgpu_apply_bcs (PGPU_BC_*)
BC_NONE, BC_PERIODIC, BC_SYMMETRY, BC_OUTFLOW, BC_INFLOW_OUTFLOW = 0, 1, 2, 3, 4


class CoulombParams(C.Structure):
    _fields_ = [("Clog", C.c_double), ("angular_scattering", C.c_int), ("NxN", C.c_int), ("NxN_Nthresh", C.c_int),
                ("num_subcycles", C.c_int), ("enforce_conservations", C.c_int), ("energy_fraction", C.c_double),
                ("energy_fraction_max", C.c_double), ("beta_weight_exponent", C.c_int),
                ("sort_weighted_particles", C.c_int), ("conservation_Nmin_save", C.c_int), ("weight_method", C.c_int),
                ("include_large_angle_scattering", C.c_int), ("test_large_angle_draw", C.c_double),
                ("test_fas_draw2", C.c_double), ("test_fas_draw3", C.c_double)]


ANG_TAKIZUKA, ANG_NANBU, ANG_BOBYLEV, ANG_NANBU_FAS, ANG_NANBU_FAS_V2, ANG_ISOTROPIC = 0, 1, 2, 3, 4, 5


class ElasticParams(C.Structure):
    _fields_ = [("const_sigma", C.c_double), ("ntab", C.c_int), ("E", C.c_void_p), ("Q", C.c_void_p),
                ("xi", C.c_void_p), ("angular_scattering", C.c_int), ("use_loglog_interp", C.c_int),
                ("weight_method", C.c_int)]


class PicardStats(C.Structure):
    _fields_ = [("num_parts_its", C.c_long), ("num_apply_its", C.c_long), ("num_unconverged", C.c_long)]


_lib = None


def load():
    """Load the CUDA library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    global LIB_PATH
    if os.environ.get("PGPU_LIB"):      # kernel A/B experiments (picnic_b200/build.py --variant)
        LIB_PATH = os.path.join(os.path.dirname(LIB_PATH), "libpicnic_gpu_%s.so" % os.environ["PGPU_LIB"])
    if not os.path.exists(LIB_PATH):
        raise ImportError("picnic_b200: %s is missing -- run `python -m picnic_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.pgpu_last_error.restype = C.c_char_p
    lib.pgpu_species_count.restype = C.c_long
    lib.pgpu_launch_count.restype = C.c_long
    vp, dbl, i32, lng = C.c_void_p, C.c_double, C.c_int, C.c_long
    sig = {
        "pgpu_init": [i32], "pgpu_finalize": [], "pgpu_set_stream": [vp], "pgpu_synchronize": [],
        "pgpu_set_exact_math": [i32], "pgpu_set_deposit_mode": [i32],
        "pgpu_grid_create": [vp, vp], "pgpu_grid_destroy": [vp],
        "pgpu_fields_set": [vp, i32, vp, vp, vp], "pgpu_fields_select": [vp, i32],
        "pgpu_host_register": [vp, C.c_size_t], "pgpu_host_unregister": [vp], "pgpu_field_bounds": [vp, i32, vp, vp],
        "pgpu_current_zero": [vp], "pgpu_current_add_species": [vp, vp], "pgpu_current_finalize": [vp],
        "pgpu_current_get": [vp, i32, vp, vp, vp], "pgpu_current_get_async": [vp, i32, vp, vp, vp],
        "pgpu_species_create": [vp, vp, vp], "pgpu_species_destroy": [vp],
        "pgpu_species_set_solver_params": [vp, i32, i32, dbl],
        "pgpu_species_upload": [vp, lng, vp, vp, vp, vp, vp, vp],
        "pgpu_species_download": [vp, vp, vp, vp, vp, vp, vp],
        "pgpu_species_count": [vp],
        "pgpu_advance_positions_explicit": [vp, dbl, i32], "pgpu_advance_positions_implicit": [vp, dbl],
        "pgpu_advance_positions_2nd_half": [vp], "pgpu_interpolate_fields_to_particles": [vp],
        "pgpu_advance_velocities": [vp, dbl, i32], "pgpu_advance_velocities_2nd_half": [vp],
        "pgpu_average_velocities": [vp], "pgpu_update_old_particle_positions": [vp],
        "pgpu_update_old_particle_velocities": [vp], "pgpu_reset_particles": [vp],
        "pgpu_species_download_fields": [vp, vp, vp], "pgpu_species_upload_fields": [vp, vp, vp],
        "pgpu_scatter_delta_u": [lng, vp, vp, vp, vp, vp, vp],
        "pgpu_fab_pack_d": [vp, i32, i32, vp, vp, vp], "pgpu_fab_unpack_d": [vp, i32, i32, vp, vp, vp, i32],
        "pgpu_wire_doubles": [vp], "pgpu_species_mark_leavers": [vp, vp],
        "pgpu_species_mark_leavers_d": [vp, vp], "pgpu_species_set_leaver_counts": [vp, vp],
        "pgpu_species_pack_leavers_d": [vp, vp], "pgpu_species_append_d": [vp, lng, vp],
        "pgpu_advance_particles": [vp, dbl], "pgpu_advance_particles_iteratively": [vp, dbl, i32, vp],
        "pgpu_set_current_density": [vp, dbl, i32], "pgpu_species_current_get": [vp, i32, vp, vp, vp],
        "pgpu_set_charge_density": [vp, vp, vp, vp, vp],
        "pgpu_explicit_step": [vp, C.c_double, vp, vp, C.c_int],
        "pgpu_species_outflow_download": [vp] * 8, "pgpu_species_outflow_fluxes": [vp, vp],
        "pgpu_remove_outflow_particles": [vp], "pgpu_species_append": [vp, C.c_long] + [vp] * 6,
        "pgpu_species_inflow_append": [vp, C.c_long] + [vp] * 4 + [i32, i32], "pgpu_species_inflow_count": [vp],
        "pgpu_species_inflow_download": [vp] * 8, "pgpu_species_inflow_clear": [vp],
        "pgpu_advance_inflow_particles_and_set_J": [vp, dbl, i32],
        "pgpu_species_inflow_current_get": [vp, i32, vp, vp, vp], "pgpu_current_add_inflow": [vp, vp],
        "pgpu_species_inflow_fluxes": [vp, vp],
        "pgpu_species_set_suborbit_model": [vp, C.c_int, C.c_int], "pgpu_transfer_fast_particles": [vp],
        "pgpu_advance_suborbit_particles_and_set_J": [vp, C.c_double, C.c_int],
        "pgpu_species_suborbit_current_get": [vp, C.c_int, vp, vp, vp], "pgpu_current_add_suborbit": [vp, vp],
        "pgpu_merge_suborbit_particles": [vp], "pgpu_species_suborbit_download": [vp] * 8,
        "pgpu_grid_set_external_fields": [vp, vp], "pgpu_add_external_fields_to_particles": [vp],
        "pgpu_fields_packed_size": [vp, vp], "pgpu_fields_set_packed": [vp, vp],
        "pgpu_current_packed_size": [vp, vp], "pgpu_current_get_packed_async": [vp, vp],
        "pgpu_bin_particles": [vp], "pgpu_sort_for_locality": [vp], "pgpu_species_cell_index": [vp, vp],
        "pgpu_species_cell_offsets": [vp, vp], "pgpu_set_moments_from_bins": [vp],
        "pgpu_species_moments_get": [vp, vp, vp, vp], "pgpu_debye_length": [vp, vp, i32, vp],
        "pgpu_apply_bcs": [vp, vp, vp], "pgpu_finish_implicit_step": [vp, vp, vp], "pgpu_stable_dt": [vp, vp], "pgpu_global_moments": [vp, vp],
        "pgpu_collide_ta": [vp, vp, dbl, dbl, C.c_uint64, C.c_uint64, vp],
        "pgpu_ta_delta_u": [lng, vp, vp, vp, vp, dbl, dbl, dbl, vp, vp, vp, vp],
        "pgpu_ta_lorentz_scatter": [lng, vp, vp, dbl, dbl, vp, dbl, dbl, dbl, vp, vp, vp, vp, vp],
        "pgpu_collide_coulomb": [vp, vp, vp, dbl, C.c_uint64, C.c_uint64, vp],
        "pgpu_coulomb_delta_u": [lng, vp, vp, dbl, dbl, dbl, dbl, vp, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "pgpu_coulomb_lorentz_scatter": [lng, vp, vp, vp, dbl, dbl, dbl, dbl, vp, dbl] + [vp] * 10,
        "pgpu_collide_elastic": [vp, vp, vp, dbl, C.c_uint64, C.c_uint64, vp],
        "pgpu_collide_hard_sphere": [vp, vp, dbl, dbl, C.c_uint64, C.c_uint64, vp],
        "pgpu_scatter_nu_max_hard_sphere": [vp, vp, dbl, vp],
        "pgpu_collide_hard_sphere_wm": [vp, vp, dbl, i32, dbl, C.c_uint64, C.c_uint64, vp],
        "pgpu_collide_vhs": [vp, dbl, dbl, dbl, dbl, C.c_uint64, C.c_uint64, vp],
        "pgpu_scatter_nu_max_vhs": [vp, dbl, dbl, dbl, vp],
        "pgpu_scatter_nu_max_ta": [vp, vp, dbl, vp], "pgpu_scatter_nu_max_coulomb": [vp, vp, vp, vp],
        "pgpu_scatter_nu_max_elastic": [vp, vp, vp, vp],
        "pgpu_halo_create": [vp, i32, vp, vp], "pgpu_halo_create_rho": [vp, vp, i32, vp, vp],
        "pgpu_current_filter": [vp, i32, i32], "pgpu_charge_density_filter": [vp, vp],
        "pgpu_charge_density_deposit": [vp, vp], "pgpu_charge_density_get": [vp, vp, vp, vp, vp], "pgpu_halo_destroy": [vp], "pgpu_halo_phases": [vp],
        "pgpu_halo_area_offset": [vp, i32, vp, vp], "pgpu_halo_inbox": [vp, vp, vp],
        "pgpu_halo_ipc_handle": [vp, vp], "pgpu_halo_ipc_open": [vp, vp, vp],
        "pgpu_halo_connect": [vp, i32, vp, i32, lng], "pgpu_halo_begin": [vp], "pgpu_halo_send": [vp, i32],
        "pgpu_halo_recv_add": [vp, i32],
        "pgpu_migrator_create": [vp, lng, vp], "pgpu_migrator_destroy": [vp], "pgpu_migrator_inbox": [vp, vp, vp],
        "pgpu_migrator_ipc_handle": [vp, vp], "pgpu_migrator_ipc_open": [vp, vp, vp],
        "pgpu_migrator_connect": [vp, i32, vp], "pgpu_migrate_send": [vp], "pgpu_migrate_recv": [vp],
        "pgpu_migrate_finish": [vp, vp, vp, vp],
        "pgpu_mass_matrices_init": [vp, i32, vp], "pgpu_mass_matrices_ncomp": [vp, vp],
        "pgpu_mass_matrices_zero": [vp], "pgpu_accumulate_mass_matrices": [vp, dbl],
        "pgpu_mass_matrices_save_E0": [vp], "pgpu_compute_J_from_mass_matrices": [vp],
        "pgpu_mass_matrix_get": [vp, i32, vp, vp, vp, i32], "pgpu_mass_matrix_J0_get": [vp, i32, vp, vp, vp],
        "pgpu_profile_enable": [i32], "pgpu_profile_reset": [], "pgpu_profile_query": [C.c_char_p, vp, vp],
        "pgpu_species_deferred_count": [vp, vp],
        "pgpu_apply_forces_curvilinear": [vp, i32, dbl, i32, i32],
        "pgpu_species_virtual_positions_set": [vp, vp], "pgpu_species_virtual_positions_get": [vp, vp],
        "pgpu_particle_linear_size": [vp], "pgpu_species_download_linear": [vp, vp],
        "pgpu_species_upload_linear": [vp, lng, vp],
        "pgpu_launch_count": [], "pgpu_picard_totals": [vp, vp, vp, i32], "pgpu_abi_version": [], "pgpu_last_error": [],
    }
    for name, args in sig.items():
        getattr(lib, name).argtypes = args
    _lib = lib
    return lib


class HaloMsg(C.Structure):
    """pgpu_halo_msg"""
    _fields_ = [("phase", C.c_int), ("recv_area", C.c_int), ("lo", (C.c_int * 2) * 3), ("hi", (C.c_int * 2) * 3)]


def check(rc):
    if rc != 0:
        raise PgpuError(rc, load().pgpu_last_error().decode())


def init(device=0):
    check(load().pgpu_init(device))


def finalize():
    load().pgpu_finalize()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i2(v):
    return (C.c_int * 2)(*(list(v) + [0, 0])[:2])


class Grid:
    def __init__(self, D, ncell, xmin, dx, nghost, periodic, box_lo=None, box_hi=None, volume_scale=1.0):
        d = GridDesc()
        d.D = D
        box_lo = [0] * D if box_lo is None else box_lo
        box_hi = [n - 1 for n in ncell] if box_hi is None else box_hi
        for k in range(D):
            d.ncell[k], d.xmin[k], d.dx[k] = ncell[k], xmin[k], dx[k]
            d.periodic[k], d.box_lo[k], d.box_hi[k] = int(periodic[k]), box_lo[k], box_hi[k]
        d.nghost = nghost
        d.volume_scale = volume_scale
        self.desc = d
        self.D = D
        self.h = C.c_void_p()
        check(load().pgpu_grid_create(C.byref(d), C.byref(self.h)))

    def field_bounds(self, comp):
        lo, hi = (C.c_int * 2)(), (C.c_int * 2)()
        check(load().pgpu_field_bounds(self.h, comp, lo, hi))
        return tuple(lo[:self.D]), tuple(hi[:self.D])

    def set_field(self, comp, arr, lo, hi):
        a = np.asfortranarray(arr, dtype=np.float64)
        check(load().pgpu_fields_set(self.h, comp, _p(a), _i2(lo), _i2(hi)))
        self._keep = a  # async H2D: keep the host buffer alive until the next sync

    def set_external_fields(self, six):
        """six: list of 6 ExtFn (Ex Ey Ez Bx By Bz) or None to switch them off."""
        if six is None:
            check(load().pgpu_grid_set_external_fields(self.h, None))
        else:
            arr = (ExtFn * 6)(*six)
            check(load().pgpu_grid_set_external_fields(self.h, arr))

    def current_filter(self, in_plane=True, virtual=True):
        """PicSpeciesInterface::filterJ's binomial filter on the summed J (after current_finalize / the halo)."""
        check(load().pgpu_current_filter(self.h, int(in_plane), int(virtual)))

    def charge_density_filter(self, stag):
        check(load().pgpu_charge_density_filter(self.h, _i2(stag)))

    def charge_density_get(self, stag):
        """The grid's resident charge-density array of one centring (after Species.charge_density_deposit and the
        ghost add-exchange) -> (array, lo, hi)."""
        g = self.desc
        D = g.D
        lo = [g.box_lo[k] - g.nghost for k in range(D)]
        hi = [g.box_hi[k] + g.nghost + stag[k] for k in range(D)]
        out = np.zeros(tuple(h - l + 1 for l, h in zip(lo, hi)), order="F")
        check(load().pgpu_charge_density_get(self.h, _i2(stag), _p(out), _i2(lo), _i2(hi)))
        return out, lo, hi

    def fields_packed_size(self):
        n = C.c_long()
        check(load().pgpu_fields_packed_size(self.h, C.byref(n)))
        return n.value

    def current_packed_size(self):
        n = C.c_long()
        check(load().pgpu_current_packed_size(self.h, C.byref(n)))
        return n.value

    def set_fields(self, E, B):
        keep = []
        for c, (lo, hi, a) in enumerate(list(E) + list(B)):
            a = np.asfortranarray(a, dtype=np.float64)
            keep.append(a)
            check(load().pgpu_fields_set(self.h, c, _p(a), _i2(lo), _i2(hi)))
        check(load().pgpu_synchronize())

    def fields_select(self, slot):
        check(load().pgpu_fields_select(self.h, slot))

    def current_zero(self):
        check(load().pgpu_current_zero(self.h))

    def current_add(self, sp):
        check(load().pgpu_current_add_species(self.h, sp.h))

    def current_add_inflow(self, sp):
        check(load().pgpu_current_add_inflow(self.h, sp.h))

    def current_finalize(self):
        check(load().pgpu_current_finalize(self.h))

    def current_get(self, comp):
        lo, hi = self.field_bounds(comp)
        shape = tuple(h - l + 1 for l, h in zip(lo, hi))
        out = np.zeros(shape, order="F")
        check(load().pgpu_current_get(self.h, comp, _p(out), _i2(lo), _i2(hi)))
        return out

    # ---- mass matrices (PicSpeciesInterface::initializeMassMatrices / setMassMatrices / computeJfromMassMatrices)
    def mass_matrices_init(self, interp):
        nc = np.zeros((9, 2), dtype=np.int32)
        check(load().pgpu_mass_matrices_init(self.h, interp, _p(nc)))
        self.mm_ncomp = nc
        return nc

    def mass_matrices_zero(self):
        check(load().pgpu_mass_matrices_zero(self.h))

    def mass_matrices_save_E0(self):
        check(load().pgpu_mass_matrices_save_E0(self.h))

    def compute_J_from_mass_matrices(self):
        check(load().pgpu_compute_J_from_mass_matrices(self.h))

    def mass_matrix_get(self, which):
        """sigma `which` (0..8 = xx xy xz yx yy yz zx zy zz) as an array [box of the row's J component] + (ncomp,)"""
        lo, hi = self.field_bounds(which // 3)
        ncomp = int(self.mm_ncomp[which, 0]) * int(self.mm_ncomp[which, 1])
        shape = tuple(h - l + 1 for l, h in zip(lo, hi)) + (ncomp,)
        out = np.zeros(shape, order="F")
        check(load().pgpu_mass_matrix_get(self.h, which, _p(out), _i2(lo), _i2(hi), ncomp))
        return out

    def mass_matrix_J0_get(self, comp):
        lo, hi = self.field_bounds(comp)
        out = np.zeros(tuple(h - l + 1 for l, h in zip(lo, hi)), order="F")
        check(load().pgpu_mass_matrix_J0_get(self.h, comp, _p(out), _i2(lo), _i2(hi)))
        return out

    def debye_length(self, species):
        arr = (C.c_void_p * len(species))(*[s.h for s in species])
        nbox = [self.desc.box_hi[k] - self.desc.box_lo[k] + 1 for k in range(self.D)]
        out = np.zeros(int(np.prod(nbox)))
        check(load().pgpu_debye_length(self.h, arr, len(species), _p(out)))
        return out

    def destroy(self):
        if self.h:
            load().pgpu_grid_destroy(self.h)
            self.h = None


class Species:
    def __init__(self, grid, mass, charge, fnorm_const, cvac_norm, interp_N=TSC, interp_J=CC1, interp_E=CC1,
                 rtol=1e-12, iter_max=21, order_swap=0, bc_check_lo=(0, 0), bc_check_hi=(0, 0),
                 motion=1, forces=1, relativistic=False, higuera_cary=False):
        d = SpeciesDesc()
        d.mass, d.charge, d.fnorm_const, d.cvac_norm = mass, charge, fnorm_const, cvac_norm
        d.interp_N, d.interp_J, d.interp_E = interp_N, interp_J, interp_E
        d.rtol, d.iter_max, d.order_swap = rtol, iter_max, int(order_swap)
        for k in range(2):
            d.bc_check_lo[k] = bc_check_lo[k] if k < len(bc_check_lo) else 0
            d.bc_check_hi[k] = bc_check_hi[k] if k < len(bc_check_hi) else 0
        d.motion, d.forces = int(motion), int(forces)
        d.relativistic, d.higuera_cary = int(relativistic), int(higuera_cary)
        self.desc = d
        self.grid = grid
        self.D = grid.D
        self.h = C.c_void_p()
        check(load().pgpu_species_create(grid.h, C.byref(d), C.byref(self.h)))

    @property
    def n(self):
        return load().pgpu_species_count(self.h)

    def upload(self, x, v, w, xold=None, vold=None, ids=None):
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        x, v, w, xold, vold = c(x), c(v), c(w), c(xold), c(vold)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        n = w.size
        assert x.shape == (self.D, n) and v.shape == (3, n)
        check(load().pgpu_species_upload(self.h, n, _p(x), _p(xold), _p(v), _p(vold), _p(w), _p(ids)))

    def download(self):
        n = self.n
        out = {"x": np.zeros((self.D, n)), "xold": np.zeros((self.D, n)), "v": np.zeros((3, n)),
               "vold": np.zeros((3, n)), "w": np.zeros(n), "id": np.zeros(n, dtype=np.uint64)}
        check(load().pgpu_species_download(self.h, _p(out["x"]), _p(out["xold"]), _p(out["v"]),
                                           _p(out["vold"]), _p(out["w"]), _p(out["id"])))
        return out

    def advance_positions_explicit(self, dt, half=False):
        check(load().pgpu_advance_positions_explicit(self.h, dt, int(half)))

    def advance_positions_implicit(self, dt):
        check(load().pgpu_advance_positions_implicit(self.h, dt))

    def advance_positions_2nd_half(self):
        check(load().pgpu_advance_positions_2nd_half(self.h))

    # ---- outflow lists / inflow injection ---------------------------------------------------------------
    @property
    def n_outflow(self):
        f = load().pgpu_species_outflow_count
        f.restype = C.c_long
        f.argtypes = [C.c_void_p]
        return f(self.h)

    def outflow_download(self):
        n = self.n_outflow
        out = {"x": np.zeros((self.D, n)), "xold": np.zeros((self.D, n)), "v": np.zeros((3, n)), "vold": np.zeros((3, n)),
               "w": np.zeros(n), "id": np.zeros(n, dtype=np.uint64), "boundary": np.zeros(n, dtype=np.int32)}
        if n:
            check(load().pgpu_species_outflow_download(self.h, _p(out["x"]), _p(out["xold"]), _p(out["v"]), _p(out["vold"]),
                                                       _p(out["w"]), _p(out["id"]), _p(out["boundary"])))
        return out

    def outflow_fluxes(self):
        out = np.zeros((4, 5))
        check(load().pgpu_species_outflow_fluxes(self.h, _p(out)))
        return out

    def remove_outflow(self):
        check(load().pgpu_remove_outflow_particles(self.h))

    def append(self, x, v, w, xold=None, vold=None, ids=None):
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        x, v, w, xold, vold = c(x), c(v), c(w), c(xold), c(vold)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        check(load().pgpu_species_append(self.h, w.size, _p(x), _p(xold), _p(v), _p(vold), _p(w), _p(ids)))

    # ---- inflow lists (suborbit_inflow_J) -------------------------------------------------------------
    def inflow_append(self, x, v, w, bdry_dir, bdry_side, ids=None):
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x, v, w = c(x), c(v), c(w)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        check(load().pgpu_species_inflow_append(self.h, w.size, _p(x), _p(v), _p(w), _p(ids), bdry_dir, bdry_side))

    @property
    def n_inflow(self):
        f = load().pgpu_species_inflow_count
        f.restype = C.c_long
        return f(self.h)

    def inflow_download(self):
        n = self.n_inflow
        out = {"x": np.zeros((self.D, n)), "xold": np.zeros((self.D, n)), "v": np.zeros((3, n)), "vold": np.zeros((3, n)),
               "w": np.zeros(n), "id": np.zeros(n, dtype=np.uint64), "code": np.zeros(n, dtype=np.int32)}
        if n:
            check(load().pgpu_species_inflow_download(self.h, _p(out["x"]), _p(out["xold"]), _p(out["v"]), _p(out["vold"]),
                                                      _p(out["w"]), _p(out["id"]), _p(out["code"])))
        out["nsub"], out["boundary"] = out["code"] >> 3, out["code"] & 7
        return out

    def advance_inflow_and_set_J(self, dt, from_emjacobian=False):
        check(load().pgpu_advance_inflow_particles_and_set_J(self.h, dt, int(from_emjacobian)))

    def inflow_current_get(self, comp):
        lo, hi = self.grid.field_bounds(comp)
        out = np.zeros(tuple(h - l + 1 for l, h in zip(lo, hi)), order="F")
        check(load().pgpu_species_inflow_current_get(self.h, comp, _p(out), _i2(lo), _i2(hi)))
        return out

    def inflow_fluxes(self):
        out = np.zeros(20)
        check(load().pgpu_species_inflow_fluxes(self.h, _p(out)))
        return out.reshape(4, 5)

    # ---- sub-orbit model --------------------------------------------------------------------------
    def set_suborbit_model(self, use=True, fast_particles=False):
        check(load().pgpu_species_set_suborbit_model(self.h, int(use), int(fast_particles)))

    @property
    def n_suborbit(self):
        f = load().pgpu_species_suborbit_count
        f.restype = C.c_long
        f.argtypes = [C.c_void_p]
        return f(self.h)

    def transfer_fast_particles(self):
        check(load().pgpu_transfer_fast_particles(self.h))

    def advance_suborbit_and_set_J(self, dt, from_emjacobian=False):
        check(load().pgpu_advance_suborbit_particles_and_set_J(self.h, dt, int(from_emjacobian)))

    def suborbit_current_get(self, comp):
        lo, hi = self.grid.field_bounds(comp)
        out = np.zeros(tuple(h - l + 1 for l, h in zip(lo, hi)), order="F")
        check(load().pgpu_species_suborbit_current_get(self.h, comp, _p(out), _i2(lo), _i2(hi)))
        return out

    def merge_suborbit(self):
        check(load().pgpu_merge_suborbit_particles(self.h))

    def suborbit_download(self):
        n = self.n_suborbit
        out = {"x": np.zeros((self.D, n)), "xold": np.zeros((self.D, n)), "v": np.zeros((3, n)), "vold": np.zeros((3, n)),
               "w": np.zeros(n), "id": np.zeros(n, dtype=np.uint64), "nsub": np.zeros(n, dtype=np.int32)}
        if n:
            check(load().pgpu_species_suborbit_download(self.h, _p(out["x"]), _p(out["xold"]), _p(out["v"]), _p(out["vold"]),
                                                        _p(out["w"]), _p(out["id"]), _p(out["nsub"])))
        return out

    def download_linear(self):
        """[n, 2 D + 10] array of JustinsParticle::linearOut records"""
        lib = load()
        lib.pgpu_particle_linear_size.restype = C.c_long
        nw = lib.pgpu_particle_linear_size(self.h) // 8
        rec = np.zeros((self.n, nw))
        check(lib.pgpu_species_download_linear(self.h, _p(rec)))
        return rec

    def upload_linear(self, rec):
        rec = np.ascontiguousarray(rec, dtype=np.float64)
        check(load().pgpu_species_upload_linear(self.h, rec.shape[0], _p(rec)))

    def apply_forces_curvilinear(self, push_type, dt, by_half, anticyclic=False):
        check(load().pgpu_apply_forces_curvilinear(self.h, push_type, dt, int(by_half), int(anticyclic)))

    def set_virtual_positions(self, virt):
        virt = np.ascontiguousarray(virt, dtype=np.float64)
        assert virt.shape == (2, self.n)
        check(load().pgpu_species_virtual_positions_set(self.h, _p(virt)))

    def virtual_positions(self):
        out = np.zeros((2, self.n))
        check(load().pgpu_species_virtual_positions_get(self.h, _p(out)))
        return out

    def deferred_count(self):
        n = C.c_long(0)
        check(load().pgpu_species_deferred_count(self.h, C.byref(n)))
        return n.value

    def explicit_step(self, dt, bc_lo, bc_hi, second_half=False):
        check(load().pgpu_explicit_step(self.h, dt, _i2(bc_lo), _i2(bc_hi), int(second_half)))

    def add_external_fields(self):
        check(load().pgpu_add_external_fields_to_particles(self.h))

    def interpolate_fields(self):
        check(load().pgpu_interpolate_fields_to_particles(self.h))

    def particle_fields(self):
        n = self.n
        Ep, Bp = np.zeros((3, n)), np.zeros((3, n))
        check(load().pgpu_species_download_fields(self.h, _p(Ep), _p(Bp)))
        return Ep, Bp

    def set_particle_fields(self, Ep, Bp):
        Ep, Bp = np.ascontiguousarray(Ep, dtype=np.float64), np.ascontiguousarray(Bp, dtype=np.float64)
        check(load().pgpu_species_upload_fields(self.h, _p(Ep), _p(Bp)))

    def advance_velocities(self, dt, half):
        check(load().pgpu_advance_velocities(self.h, dt, int(half)))

    def advance_velocities_2nd_half(self):
        check(load().pgpu_advance_velocities_2nd_half(self.h))

    def average_velocities(self):
        check(load().pgpu_average_velocities(self.h))

    def update_old_positions(self):
        check(load().pgpu_update_old_particle_positions(self.h))

    def update_old_velocities(self):
        check(load().pgpu_update_old_particle_velocities(self.h))

    def reset_particles(self):
        check(load().pgpu_reset_particles(self.h))

    def advance_particles(self, dt):
        check(load().pgpu_advance_particles(self.h, dt))
        check(load().pgpu_synchronize())

    def advance_iteratively(self, dt, deposit=False, stats=True):
        st = PicardStats()
        check(load().pgpu_advance_particles_iteratively(self.h, dt, int(deposit), C.byref(st) if stats else None))
        return st if stats else None

    def accumulate_mass_matrices(self, dt):
        check(load().pgpu_accumulate_mass_matrices(self.h, dt))

    def set_current_density(self, dt, from_explicit=False):
        check(load().pgpu_set_current_density(self.h, dt, int(from_explicit)))

    def current_get(self, comp):
        lo, hi = self.grid.field_bounds(comp)
        shape = tuple(h - l + 1 for l, h in zip(lo, hi))
        out = np.zeros(shape, order="F")
        check(load().pgpu_species_current_get(self.h, comp, _p(out), _i2(lo), _i2(hi)))
        return out

    def charge_density(self, stag):
        g = self.grid.desc
        lo = [g.box_lo[k] - g.nghost for k in range(self.D)]
        hi = [g.box_hi[k] + g.nghost + stag[k] for k in range(self.D)]
        out = np.zeros(tuple(h - l + 1 for l, h in zip(lo, hi)), order="F")
        check(load().pgpu_set_charge_density(self.h, _i2(stag), _p(out), _i2(lo), _i2(hi)))
        return out, lo, hi

    def charge_density_deposit(self, stag):
        """First half of charge_density for a domain of several boxes: the scaled deposit stays in the grid's
        resident array of this centring; a PeerHaloExchange(..., rho_stag=stag) then adds the ghost layers of
        neighbouring boxes, and Grid.charge_density_get reads the array."""
        check(load().pgpu_charge_density_deposit(self.h, _i2(stag)))

    def bin_particles(self):
        check(load().pgpu_bin_particles(self.h))

    def sort_for_locality(self):
        check(load().pgpu_sort_for_locality(self.h))

    def cell_index(self):
        out = np.zeros((self.D, self.n), dtype=np.int32)
        check(load().pgpu_species_cell_index(self.h, _p(out)))
        return out

    def cell_offsets(self):
        g = self.grid.desc
        ncell = int(np.prod([g.box_hi[k] - g.box_lo[k] + 1 for k in range(self.D)]))
        out = np.zeros(ncell + 1, dtype=np.int64)
        check(load().pgpu_species_cell_offsets(self.h, _p(out)))
        return out

    def set_moments(self):
        check(load().pgpu_set_moments_from_bins(self.h))

    def moments(self):
        g = self.grid.desc
        ncell = int(np.prod([g.box_hi[k] - g.box_lo[k] + 1 for k in range(self.D)]))
        dens, mom, ene = np.zeros(ncell), np.zeros((3, ncell)), np.zeros((3, ncell))
        check(load().pgpu_species_moments_get(self.h, _p(dens), _p(mom), _p(ene)))
        return dens, mom, ene

    def apply_bcs(self, bc_lo, bc_hi):
        check(load().pgpu_apply_bcs(self.h, _i2(bc_lo), _i2(bc_hi)))

    def finish_implicit_step(self, bc_lo, bc_hi):
        check(load().pgpu_finish_implicit_step(self.h, _i2(bc_lo), _i2(bc_hi)))

    def stable_dt(self):
        out = C.c_double(0)
        check(load().pgpu_stable_dt(self.h, C.byref(out)))
        return out.value

    def global_moments(self):
        out = np.zeros(7)
        check(load().pgpu_global_moments(self.h, _p(out)))
        return out

    def destroy(self):
        if self.h:
            load().pgpu_species_destroy(self.h)
            self.h = None


def collide_ta(sA, sB, Clog, dt_sec, seed, step, count=True):
    np_ = C.c_long(0)
    check(load().pgpu_collide_ta(sA.h, sB.h, Clog, dt_sec, seed, step, C.byref(np_) if count else None))
    return np_.value


def ta_delta_u(vp1, den1, vp2, den2, b90_fact, Clog, dt_sec, gauss, u_theta, u_phi):
    n = den1.size
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    out = np.zeros((3, n))
    args = [c(vp1), c(den1), c(vp2), c(den2)]
    rnd = [c(gauss), c(u_theta), c(u_phi)]
    check(load().pgpu_ta_delta_u(n, _p(args[0]), _p(args[1]), _p(args[2]), _p(args[3]), b90_fact, Clog, dt_sec,
                                 _p(rnd[0]), _p(rnd[1]), _p(rnd[2]), _p(out)))
    return out


def collide_coulomb(sA, sB, Clog, dt_sec, seed, step, angular=0, NxN=False, NxN_Nthresh=11, num_subcycles=1,
                    count=True, enforce=False, energy_fraction=0.05, energy_fraction_max=0.5, beta_weight_exponent=1,
                    conservative=False, large_angle=False):
    prm = CoulombParams(Clog, angular, int(NxN), NxN_Nthresh, num_subcycles, int(enforce), energy_fraction,
                        energy_fraction_max, beta_weight_exponent, 0, 0, int(conservative), int(large_angle), 0.5)
    np_ = C.c_long(0)
    check(load().pgpu_collide_coulomb(sA.h, sB.h, C.byref(prm), dt_sec, seed, step, C.byref(np_) if count else None))
    return np_.value


def collide_hard_sphere(sA, sB, sigmaT, dt_sec, seed, step, count=True):
    np_ = C.c_long(0)
    check(load().pgpu_collide_hard_sphere(sA.h, sB.h, sigmaT, dt_sec, seed, step, C.byref(np_) if count else None))
    return np_.value


def collide_hard_sphere_conservative(sp, sigmaT, dt_sec, seed, step, sp2=None):
    """weight_method CONSERVATIVE: self-scattering of sp, or sp against sp2"""
    np_ = C.c_long(0)
    check(load().pgpu_collide_hard_sphere_wm(sp.h, (sp2 or sp).h, sigmaT, 1, dt_sec, seed, step, C.byref(np_)))
    return np_.value


def collide_vhs(sp, eta, T0, mu0, dt_sec, seed, step, count=True):
    np_ = C.c_long(0)
    check(load().pgpu_collide_vhs(sp.h, eta, T0, mu0, dt_sec, seed, step, C.byref(np_) if count else None))
    return np_.value


def nu_max_vhs(sp, eta, T0, mu0):
    out = C.c_double(0)
    check(load().pgpu_scatter_nu_max_vhs(sp.h, eta, T0, mu0, C.byref(out)))
    return out.value


def nu_max_hard_sphere(sA, sB, sigmaT):
    out = C.c_double(0)
    check(load().pgpu_scatter_nu_max_hard_sphere(sA.h, sB.h, sigmaT, C.byref(out)))
    return out.value


def coulomb_lorentz_scatter(up1, up2, scatter2, q1, q2, m1, m2, Clog, angular, dt_sec, EF_norm, den12, bmax, sigma_max,
                            gauss, upol, uphi, large_angle=None, fas_draws=(0.5, 0.5)):
    """Coulomb::LorentzScatter for n pairs ([3][n] arrays): (up1', up2', s12).  large_angle = the uniform draw of the
    large-angle event (include_large_angle_scattering on), None = off; fas_draws = the second and third uniform of
    NANBU_FAS(_v2)."""
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    n = np.asarray(den12).size
    prm = CoulombParams(Clog, angular, 0, 11, 1)
    prm.test_fas_draw2, prm.test_fas_draw3 = float(fas_draws[0]), float(fas_draws[1])
    if large_angle is not None:
        prm.include_large_angle_scattering, prm.test_large_angle_draw = 1, float(large_angle)
    a = [c(up1), c(up2), c(EF_norm), c(den12), c(bmax), c(sigma_max), c(gauss), c(upol), c(uphi)]
    s2 = np.ascontiguousarray(scatter2, dtype=np.int32)
    o1, o2, s12 = np.zeros((3, n)), np.zeros((3, n)), np.zeros(n)
    check(load().pgpu_coulomb_lorentz_scatter(n, _p(a[0]), _p(a[1]), _p(s2), q1, q2, m1, m2, C.byref(prm), dt_sec,
                                              *[_p(x) for x in a[2:]], _p(o1), _p(o2), _p(s12)))
    return o1, o2, s12


def coulomb_delta_u(vp1, vp2, q1, q2, m1, m2, Clog, angular, dt_sec, EF_norm, den12, bmax, sigma_max, gauss, upol, uphi,
                    large_angle=None, fas_draws=(0.5, 0.5)):
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    n = np.asarray(den12).size
    prm = CoulombParams(Clog, angular, 0, 11, 1)
    prm.test_fas_draw2, prm.test_fas_draw3 = float(fas_draws[0]), float(fas_draws[1])
    if large_angle is not None:
        prm.include_large_angle_scattering, prm.test_large_angle_draw = 1, float(large_angle)
    a = [c(vp1), c(vp2), c(EF_norm), c(den12), c(bmax), c(sigma_max), c(gauss), c(upol), c(uphi)]
    dU, s12 = np.zeros((3, n)), np.zeros(n)
    check(load().pgpu_coulomb_delta_u(n, _p(a[0]), _p(a[1]), q1, q2, m1, m2, C.byref(prm), dt_sec, *[_p(x) for x in a[2:]],
                                      _p(dU), _p(s12)))
    return dU, s12


def collide_elastic(sA, sB, dt_sec, seed, step, const_sigma=0.0, E=None, Q=None, xi=None, angular=0, loglog=False,
                    count=True, conservative=False):
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    E, Q, xi = c(E), c(Q), c(xi)
    prm = ElasticParams(const_sigma, 0 if E is None else E.size, _p(E), _p(Q), _p(xi), angular, int(loglog),
                        int(conservative))
    nc = C.c_long(0)
    check(load().pgpu_collide_elastic(sA.h, sB.h, C.byref(prm), dt_sec, seed, step, C.byref(nc) if count else None))
    return nc.value


def nu_max_ta(sA, sB, Clog):
    """TakizukaAbe::setMeanFreeTime: box maximum of the collision frequency [Hz] (scatter dt = 1/nu_max)."""
    out = C.c_double(0.0)
    check(load().pgpu_scatter_nu_max_ta(sA.h, sB.h, Clog, C.byref(out)))
    return out.value


def nu_max_coulomb(sA, sB, Clog, angular=0):
    prm = CoulombParams(Clog, angular, 0, 11, 1)
    out = C.c_double(0.0)
    check(load().pgpu_scatter_nu_max_coulomb(sA.h, sB.h, C.byref(prm), C.byref(out)))
    return out.value


def nu_max_elastic(sA, sB, const_sigma=0.0, E=None, Q=None, xi=None, angular=0, loglog=False):
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    E, Q, xi = c(E), c(Q), c(xi)
    prm = ElasticParams(const_sigma, 0 if E is None else E.size, _p(E), _p(Q), _p(xi), angular, int(loglog))
    out = C.c_double(0.0)
    check(load().pgpu_scatter_nu_max_elastic(sA.h, sB.h, C.byref(prm), C.byref(out)))
    return out.value


def scatter_delta_u(u, costh, sinth, cosphi, sinphi):
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    u = c(u)
    n = u.shape[1]
    a = [c(costh), c(sinth), c(cosphi), c(sinphi)]
    out = np.zeros((3, n))
    check(load().pgpu_scatter_delta_u(n, _p(u), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(out)))
    return out


def profile_enable(on=True):
    load().pgpu_profile_enable(int(on))


def profile_reset():
    load().pgpu_profile_reset()


def profile_query(prefix=""):
    ms, k = C.c_double(0), C.c_long(0)
    load().pgpu_profile_query(prefix.encode(), C.byref(ms), C.byref(k))
    return ms.value, k.value


def picard_totals(reset=True):
    a, b, u = C.c_long(0), C.c_long(0), C.c_long(0)
    check(load().pgpu_picard_totals(C.byref(a), C.byref(b), C.byref(u), int(reset)))
    return a.value, b.value, u.value
