"""ctypes mirror of include/crt1d_b200.h (structs, constants, function prototypes)."""
import ctypes as C

ABI_VERSION = 4

OK = 0
ERR_INVALID_ARG = -1
ERR_NULL_POINTER = -2
ERR_UNSUPPORTED = -3
ERR_CUDA = -4
ERR_NO_DEVICE = -5
ERR_NO_MEMORY = -6
NONFINITE = 1  # positive finding of crt1d_solve_host: results delivered, some scenario is non-finite
STATUS_NONFINITE = 1

SCHEME_IDS = {"2s": 0, "4s": 1, "bf": 2, "bl": 3, "g77": 4, "n79": 5, "zq": 6, "zq_pa": 7}

_pd = C.c_void_p  # double* / int32_t* carried as raw addresses (device or host)


class Batch(C.Structure):
    """struct crt1d_batch"""

    _fields_ = [
        ("n_scen", C.c_int64),
        ("n_z", C.c_int32),
        ("n_wl", C.c_int32),
        ("n_lai", C.c_int32),
        ("n_leaf", C.c_int32),
        ("n_soil", C.c_int32),
        ("n_sky", C.c_int32),
        ("psi", _pd),
        ("K_b", _pd),
        ("G", _pd),
        ("mu_bar", _pd),
        ("G_int", _pd),
        ("tau_i", _pd),
        ("tau_psi", _pd),
        ("lai_idx", _pd),
        ("leaf_idx", _pd),
        ("soil_idx", _pd),
        ("sky_idx", _pd),
        ("lai_lib", _pd),
        ("tau_d_lev", _pd),
        ("leaf_r_lib", _pd),
        ("leaf_t_lib", _pd),
        ("soil_r_lib", _pd),
        ("I_dr0_lib", _pd),
        ("I_df0_lib", _pd),
        ("mla_deg", C.c_double),
        ("mu_s", C.c_double),
    ]


class Out(C.Structure):
    """struct crt1d_out"""

    _fields_ = [
        ("I_dr", _pd),
        ("I_df_d", _pd),
        ("I_df_u", _pd),
        ("F", _pd),
        ("x0", _pd),
        ("x1", _pd),
        ("x2", _pd),
        ("rho_c", _pd),
        ("band_w", _pd),
        ("n_bw", C.c_int32),
        ("absorbed", _pd),
        ("profile_f32", C.c_int32),
        ("status", _pd),
    ]


class AbsorptionOut(C.Structure):
    """struct crt1d_absorption_out"""

    _fields_ = [(k, _pd) for k in ("aI", "aI_df", "aI_dr", "aI_sh", "aI_sl", "aI_df_sl", "aI_df_sh")]


# every symbol include/crt1d_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "crt1d_abi_version": (C.c_int, []),
    "crt1d_strerror": (C.c_char_p, [C.c_int]),
    "crt1d_last_error": (C.c_char_p, []),
    "crt1d_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "crt1d_solve": (C.c_int, [C.c_int, C.POINTER(Batch), C.POINTER(Out), C.c_void_p]),
    "crt1d_solve_2s": (C.c_int, [C.POINTER(Batch), C.POINTER(Out), C.c_void_p]),
    "crt1d_solve_4s": (C.c_int, [C.POINTER(Batch), C.POINTER(Out), C.c_void_p]),
    "crt1d_solve_bf": (C.c_int, [C.POINTER(Batch), C.POINTER(Out), C.c_void_p]),
    "crt1d_solve_bl": (C.c_int, [C.POINTER(Batch), C.POINTER(Out), C.c_void_p]),
    "crt1d_solve_g77": (C.c_int, [C.POINTER(Batch), C.POINTER(Out), C.c_void_p]),
    "crt1d_solve_n79": (C.c_int, [C.POINTER(Batch), C.POINTER(Out), C.c_void_p]),
    "crt1d_solve_zq": (C.c_int, [C.POINTER(Batch), C.POINTER(Out), C.c_void_p]),
    "crt1d_solve_zq_pa": (C.c_int, [C.POINTER(Batch), C.POINTER(Out), C.c_void_p]),
    "crt1d_solve_host": (C.c_int, [C.c_int, C.POINTER(Batch), C.POINTER(Out), C.c_int]),
    "crt1d_release_workspace": (C.c_int, []),
    "crt1d_reload_tuning": (C.c_int, []),
    "crt1d_preferred_batch": (C.c_int64, [C.c_int, C.c_int32, C.c_int32, C.c_int64, C.c_int]),
    "crt1d_calc_absorption": (
        C.c_int, [C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(AbsorptionOut), C.c_void_p]),
    "crt1d_energy_balance": (
        C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                  C.c_void_p]),
    "crt1d_leaf_G": (C.c_int, [C.c_int, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crt1d_tau_d": (C.c_int, [C.c_int, C.c_double, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crt1d_smear_tuv": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crt1d_leaf_integrals": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
}


def declare(lib):
    """Attach restype/argtypes to every exported symbol; raises AttributeError if one is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
