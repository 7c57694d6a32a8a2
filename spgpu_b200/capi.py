"""ctypes binding of the spGPU C ABI (include/spgpu.h, include/spgpu_ext.h).

The same binding class loads either OUR library (spgpu_b200/lib/libspgpu.so) or,
in tests only, the reference library built by oracle/Makefile
(oracle/_ref/libspgpu_ref.so): both export the symbols of reference
src/core/{core,ell,hell,dia,hdia,vector,*_conv}.h, so the parity tests drive both
through identical calls.  Pointers are passed as integers (device pointers come
from torch tensors' data_ptr(), host pointers from numpy arrays).

There is no CPU fallback: if the shared library is missing the import fails.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libspgpu.so")

SPGPU_SUCCESS, SPGPU_UNSUPPORTED, SPGPU_UNSPECIFIED, SPGPU_OUTOFMEMORY = 0, 1, 2, 3
SPGPU_TYPE_INT, SPGPU_TYPE_FLOAT, SPGPU_TYPE_DOUBLE = 0, 1, 2
SPGPU_TYPE_COMPLEX_FLOAT, SPGPU_TYPE_COMPLEX_DOUBLE = 3, 4


class cuFloatComplex(ctypes.Structure):
    _fields_ = [("x", c_float), ("y", c_float)]


class cuDoubleComplex(ctypes.Structure):
    _fields_ = [("x", c_double), ("y", c_double)]


class SpgpuHandleStruct(ctypes.Structure):
    """Public handle layout, reference core.h:60-82 (field order is ABI)."""
    _fields_ = [
        ("currentStream", c_void_p), ("defaultStream", c_void_p),
        ("device", c_int), ("warpSize", c_int), ("maxThreadsPerBlock", c_int),
        ("maxGridSizeX", c_int), ("maxGridSizeY", c_int), ("maxGridSizeZ", c_int),
        ("multiProcessorCount", c_int), ("capabilityMajor", c_int),
        ("capabilityMinor", c_int),
    ]


class HaloLinks(ctypes.Structure):
    """spgpuHaloLinks (include/spgpu_ext.h): where a rank's fused SpMV + halo kernels find the neighbours."""
    _fields_ = [
        ("peerLoUpperZone", c_void_p), ("peerHiLowerZone", c_void_p),
        ("myFlags", c_void_p), ("peerFlagsLo", c_void_p), ("peerFlagsHi", c_void_p),
    ]


class PeerAllreduceArgs(ctypes.Structure):
    """spgpuPeerAllreduce (include/spgpu_ext.h)."""
    _fields_ = [("world", c_int), ("myRank", c_int), ("tables", ctypes.POINTER(c_void_p)), ("seq", ctypes.c_uint)]


AR_SLOT_BYTES = 32
HALO_FLAG_WORDS = 16


class TypeInfo:
    def __init__(self, sym, ctype, rtype, np_dtype, code, is_complex):
        self.sym, self.ctype, self.rtype = sym, ctype, rtype
        self.np_dtype, self.code, self.is_complex = np.dtype(np_dtype), code, is_complex

    def scalar(self, v):
        """Python number -> by-value C argument of this type."""
        if self.is_complex:
            v = complex(v)
            return self.ctype(v.real, v.imag)
        return self.ctype(v)

    def from_c(self, v):
        if self.is_complex:
            return complex(v.x, v.y)
        return v


TYPES = {
    "I": TypeInfo("I", c_int, c_int, np.int32, SPGPU_TYPE_INT, False),
    "S": TypeInfo("S", c_float, c_float, np.float32, SPGPU_TYPE_FLOAT, False),
    "D": TypeInfo("D", c_double, c_double, np.float64, SPGPU_TYPE_DOUBLE, False),
    "C": TypeInfo("C", cuFloatComplex, c_float, np.complex64, SPGPU_TYPE_COMPLEX_FLOAT, True),
    "Z": TypeInfo("Z", cuDoubleComplex, c_double, np.complex128, SPGPU_TYPE_COMPLEX_DOUBLE, True),
}
FLOAT_SYMS = "SDCZ"
ALL_SYMS = "ISDCZ"

P = c_void_p  # every array argument


def _sig(dll, name, restype, argtypes, optional=False):
    try:
        fn = getattr(dll, name)
    except AttributeError:
        if optional:
            return None
        raise
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


def abi_symbols():
    """Every symbol include/spgpu.h declares (the reference ABI)."""
    names = ["spgpuCreate", "spgpuDestroy", "spgpuStreamCreate", "spgpuStreamDestroy",
             "spgpuSetStream", "spgpuGetStream", "spgpuSizeOf"]
    for s in FLOAT_SYMS:
        names += [f"spgpu{s}{f}spmv" for f in ("ell", "hell", "dia", "hdia")]
        names += [f"spgpu{s}ellcsput"]
        names += [f"spgpu{s}{op}" for op in (
            "dot", "mdot", "nrm2", "mnrm2", "scal", "axpby", "maxpby", "abs", "axy",
            "axypbz", "maxy", "maxypbz", "asum", "amax", "masum", "mamax")]
    for s in ALL_SYMS:
        names += [f"spgpu{s}gath", f"spgpu{s}scat", f"spgpu{s}setscal"]
    names += ["computeEllRowLenghts", "computeEllAllocPitch", "cooToEll", "ellToOell",
              "computeHellAllocSize", "ellToHell", "computeDiaDiagonalsCount", "coo2dia",
              "computeDiaAllocPitch", "getHdiaHacksCount", "computeHdiaHackOffsets",
              "diaToHdia", "computeHdiaHackOffsetsFromCoo", "cooToHdia", "bcooToBhdia", "computeBcooSize", "cooToBcoo"]
    return names


def ext_symbols():
    """Every symbol include/spgpu_ext.h declares (additive API)."""
    names = ["spgpuB200Version", "spgpuGetLaunchCount", "spgpuSetTuning", "spgpuGetTuning",
             "spgpuGetDeviceStatus", "spgpuReserveScratch", "spgpuHaloBlockPlan", "spgpuHaloBlockOf"]
    for s in FLOAT_SYMS:
        names += [f"spgpu{s}dotDev", f"spgpu{s}nrm2sqDev", f"spgpu{s}sumDev", f"spgpu{s}allreduceSumDev",
                  f"spgpu{s}axpbyDev", f"spgpu{s}hellspmvDot", f"spgpu{s}cgUpdateDev",
                  f"spgpu{s}hellspmvHalo", f"spgpu{s}hellspmvHaloDot", f"spgpu{s}hdiaspmvHalo", f"spgpu{s}hdiaspmvHaloDot"]
    names += ["spgpuCsrToHellLayoutDevice"] + [f"spgpu{s}csrToHellDevice" for s in FLOAT_SYMS]
    names += ["spgpuCsrToOhellLayoutDevice"] + [f"spgpu{s}csrToOhellDevice" for s in FLOAT_SYMS]
    names += ["spgpuHdiaHackOffsetsFromCooDevice"] + [f"spgpu{s}cooToHdiaDevice" for s in FLOAT_SYMS]
    names += ["spgpuIpcGetHandle", "spgpuIpcOpenHandle", "spgpuIpcCloseHandle",
              "spgpuDeviceAlloc", "spgpuDeviceFree", "spgpuHaloPush", "spgpuDhaloPush", "spgpuWaitFlag",
              "spgpuHaloExchange", "spgpuDhaloExchange", "spgpuHaloAck", "spgpuSetSeqCounters", "spgpuHaloSeqAdvance",
              "spgpuHaloTraceRead", "spgpuPreloadHaloKernels", "spgpuPreloadKrylovKernels"]
    return names


def mg_symbols():
    """Every symbol include/spgpu_mg.h declares (the C-level multi-GPU API)."""
    names = ["spgpuMgCreate", "spgpuMgDestroy", "spgpuMgWorld", "spgpuMgRankHandle", "spgpuMgSetExchange", "spgpuMgExchange",
             "spgpuMgSynchronize", "spgpuMgHellPlan", "spgpuMgHdiaPlan", "spgpuMgHellCreateFromBlocks", "spgpuMgMatrixDestroy", "spgpuMgMatrixHalo",
             "spgpuMgMatrixRows", "spgpuMgMatrixRowBlock", "spgpuMgVectorCreate", "spgpuMgVectorDestroy", "spgpuMgVectorSet",
             "spgpuMgVectorGet", "spgpuMgVectorLocal", "spgpuMgDcgCreate", "spgpuMgDcgStart", "spgpuMgDcgStep",
             "spgpuMgDcgSolution", "spgpuMgDcgDestroy"]
    for s in FLOAT_SYMS:
        names += [f"spgpuMg{s}hellCreate", f"spgpuMg{s}hdiaCreate", f"spgpuMg{s}hellspmv", f"spgpuMg{s}hdiaspmv",
                  f"spgpuMg{s}spmv", f"spgpuMg{s}dot", f"spgpuMg{s}nrm2", f"spgpuMg{s}axpby"]
    return names


MG_AUTO, MG_FUSED, MG_EVENTS = 0, 1, 2


class SpgpuLib:
    """One loaded spGPU-ABI shared library with typed entry points."""

    def __init__(self, path: str = LIB_PATH, ext: bool = True):
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        self.path = path
        self.dll = ctypes.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        d = self.dll
        H = c_void_p
        self.fn = {}
        f = self.fn

        # --- core
        f["spgpuCreate"] = _sig(d, "spgpuCreate", c_int, [ctypes.POINTER(c_void_p), c_int])
        f["spgpuDestroy"] = _sig(d, "spgpuDestroy", None, [H])
        f["spgpuStreamCreate"] = _sig(d, "spgpuStreamCreate", None, [H, ctypes.POINTER(c_void_p)])
        f["spgpuStreamDestroy"] = _sig(d, "spgpuStreamDestroy", None, [c_void_p])
        f["spgpuSetStream"] = _sig(d, "spgpuSetStream", None, [H, c_void_p])
        f["spgpuGetStream"] = _sig(d, "spgpuGetStream", c_void_p, [H])
        f["spgpuSizeOf"] = _sig(d, "spgpuSizeOf", c_size_t, [c_int])

        # --- SpMV + BLAS-1, per value type
        for s in FLOAT_SYMS:
            t = TYPES[s]
            T, R = t.ctype, t.rtype
            f[f"spgpu{s}ellspmv"] = _sig(d, f"spgpu{s}ellspmv", None,
                [H, P, P, T, P, P, c_int, c_int, P, P, c_int, c_int, c_int, P, T, c_int])
            f[f"spgpu{s}hellspmv"] = _sig(d, f"spgpu{s}hellspmv", None,
                [H, P, P, T, P, P, c_int, P, P, P, c_int, c_int, P, T, c_int])
            f[f"spgpu{s}diaspmv"] = _sig(d, f"spgpu{s}diaspmv", None,
                [H, P, P, T, P, P, c_int, c_int, c_int, c_int, P, T])
            f[f"spgpu{s}hdiaspmv"] = _sig(d, f"spgpu{s}hdiaspmv", None,
                [H, P, P, T, P, P, c_int, P, c_int, c_int, P, T])
            f[f"spgpu{s}ellcsput"] = _sig(d, f"spgpu{s}ellcsput", None,
                [H, T, P, P, c_int, c_int, P, c_int, P, P, P, c_int], optional=True)
            f[f"spgpu{s}dot"] = _sig(d, f"spgpu{s}dot", T, [H, c_int, P, P])
            f[f"spgpu{s}mdot"] = _sig(d, f"spgpu{s}mdot", None, [H, P, c_int, P, P, c_int, c_int])
            f[f"spgpu{s}nrm2"] = _sig(d, f"spgpu{s}nrm2", R, [H, c_int, P])
            f[f"spgpu{s}mnrm2"] = _sig(d, f"spgpu{s}mnrm2", None, [H, P, c_int, P, c_int, c_int])
            f[f"spgpu{s}scal"] = _sig(d, f"spgpu{s}scal", None, [H, P, c_int, T, P])
            f[f"spgpu{s}axpby"] = _sig(d, f"spgpu{s}axpby", None, [H, P, c_int, T, P, T, P])
            f[f"spgpu{s}maxpby"] = _sig(d, f"spgpu{s}maxpby", None,
                [H, P, c_int, T, P, T, P, c_int, c_int])
            f[f"spgpu{s}abs"] = _sig(d, f"spgpu{s}abs", None, [H, P, c_int, T, P], optional=True)
            f[f"spgpu{s}axy"] = _sig(d, f"spgpu{s}axy", None, [H, P, c_int, T, P, P], optional=True)
            f[f"spgpu{s}axypbz"] = _sig(d, f"spgpu{s}axypbz", None,
                [H, P, c_int, T, P, T, P, P], optional=True)
            f[f"spgpu{s}maxy"] = _sig(d, f"spgpu{s}maxy", None,
                [H, P, c_int, T, P, P, c_int, c_int], optional=True)
            f[f"spgpu{s}maxypbz"] = _sig(d, f"spgpu{s}maxypbz", None,
                [H, P, c_int, T, P, T, P, P, c_int, c_int], optional=True)
            f[f"spgpu{s}asum"] = _sig(d, f"spgpu{s}asum", R, [H, c_int, P])
            f[f"spgpu{s}amax"] = _sig(d, f"spgpu{s}amax", R, [H, c_int, P])
            f[f"spgpu{s}masum"] = _sig(d, f"spgpu{s}masum", None, [H, P, c_int, P, c_int, c_int])
            f[f"spgpu{s}mamax"] = _sig(d, f"spgpu{s}mamax", None, [H, P, c_int, P, c_int, c_int])
        for s in ALL_SYMS:
            T = TYPES[s].ctype
            f[f"spgpu{s}gath"] = _sig(d, f"spgpu{s}gath", None, [H, P, c_int, P, c_int, P])
            f[f"spgpu{s}scat"] = _sig(d, f"spgpu{s}scat", None, [H, P, c_int, P, P, c_int, T])
            f[f"spgpu{s}setscal"] = _sig(d, f"spgpu{s}setscal", None,
                [H, c_int, c_int, c_int, T, P], optional=True)

        # --- host conversions
        f["computeEllRowLenghts"] = _sig(d, "computeEllRowLenghts", None,
            [P, ctypes.POINTER(c_int), c_int, c_int, P, c_int])
        f["computeEllAllocPitch"] = _sig(d, "computeEllAllocPitch", c_int, [c_int])
        f["cooToEll"] = _sig(d, "cooToEll", None,
            [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, c_int, c_int])
        f["ellToOell"] = _sig(d, "ellToOell", None,
            [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int])
        f["computeHellAllocSize"] = _sig(d, "computeHellAllocSize", None,
            [ctypes.POINTER(c_int), c_int, c_int, P])
        f["ellToHell"] = _sig(d, "ellToHell", None,
            [P, P, P, c_int, P, P, c_int, c_int, P, c_int, c_int])
        f["computeDiaDiagonalsCount"] = _sig(d, "computeDiaDiagonalsCount", c_int,
            [c_int, c_int, c_int, P, P])
        f["coo2dia"] = _sig(d, "coo2dia", None,
            [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, c_int, c_int])
        f["computeDiaAllocPitch"] = _sig(d, "computeDiaAllocPitch", c_int, [c_int])
        f["getHdiaHacksCount"] = _sig(d, "getHdiaHacksCount", c_int, [c_int, c_int])
        f["computeHdiaHackOffsets"] = _sig(d, "computeHdiaHackOffsets", None,
            [ctypes.POINTER(c_int), P, c_int, P, c_int, c_int, c_int, c_int])
        f["diaToHdia"] = _sig(d, "diaToHdia", None,
            [P, P, P, c_int, P, P, c_int, c_int, c_int, c_int])
        f["computeHdiaHackOffsetsFromCoo"] = _sig(d, "computeHdiaHackOffsetsFromCoo", None,
            [ctypes.POINTER(c_int), P, c_int, c_int, c_int, c_int, P, P, c_int])
        f["cooToHdia"] = _sig(d, "cooToHdia", None,
            [P, P, P, c_int, c_int, c_int, c_int, P, P, P, c_int, c_int])

        f["bcooToBhdia"] = _sig(d, "bcooToBhdia", None,
            [P, P, P, c_int, c_int, c_int, c_int, P, P, P, c_int, c_int, c_int])
        f["computeBcooSize"] = _sig(d, "computeBcooSize", c_int, [c_int, c_int, P, P, c_int])
        f["cooToBcoo"] = _sig(d, "cooToBcoo", None, [P, P, P, c_int, c_int, P, P, P, c_int, c_int])

        # --- additive API (ours only)
        self.has_ext = False
        if ext and hasattr(d, "spgpuB200Version"):
            self.has_ext = True
            f["spgpuB200Version"] = _sig(d, "spgpuB200Version", ctypes.c_char_p, [])
            f["spgpuGetLaunchCount"] = _sig(d, "spgpuGetLaunchCount", ctypes.c_ulonglong, [H])
            f["spgpuSetTuning"] = _sig(d, "spgpuSetTuning", c_int, [H, ctypes.c_char_p, c_int])
            f["spgpuGetTuning"] = _sig(d, "spgpuGetTuning", c_int, [H, ctypes.c_char_p])
            for s in FLOAT_SYMS:
                f[f"spgpu{s}dotDev"] = _sig(d, f"spgpu{s}dotDev", None, [H, c_int, P, P, P], optional=True)
                f[f"spgpu{s}nrm2sqDev"] = _sig(d, f"spgpu{s}nrm2sqDev", None, [H, c_int, P, P], optional=True)
            f["spgpuGetDeviceStatus"] = _sig(d, "spgpuGetDeviceStatus", c_int, [H, c_int], optional=True)
            f["spgpuReserveScratch"] = _sig(d, "spgpuReserveScratch", c_int, [H, c_size_t], optional=True)
            f["spgpuHaloBlockPlan"] = _sig(d, "spgpuHaloBlockPlan", c_int,
                [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_uint)], optional=True)
            f["spgpuHaloBlockOf"] = _sig(d, "spgpuHaloBlockOf", ctypes.c_uint, [ctypes.POINTER(ctypes.c_uint), ctypes.c_uint],
                                         optional=True)
            LK, AR, U = ctypes.POINTER(HaloLinks), ctypes.POINTER(PeerAllreduceArgs), ctypes.c_uint
            for s in FLOAT_SYMS:
                T = TYPES[s].ctype
                f[f"spgpu{s}sumDev"] = _sig(d, f"spgpu{s}sumDev", None, [H, c_int, P, P], optional=True)
                f[f"spgpu{s}allreduceSumDev"] = _sig(d, f"spgpu{s}allreduceSumDev", None, [H, P, AR], optional=True)
                f[f"spgpu{s}axpbyDev"] = _sig(d, f"spgpu{s}axpbyDev", None,
                    [H, P, c_int, P, P, c_double, P, P, P, c_double, P], optional=True)
                f[f"spgpu{s}hellspmvDot"] = _sig(d, f"spgpu{s}hellspmvDot", None,
                    [H, P, P, P, c_int, P, P, c_int, P, c_int, c_int, P], optional=True)
                f[f"spgpu{s}cgUpdateDev"] = _sig(d, f"spgpu{s}cgUpdateDev", None, [H, P, P, P, P, c_int, P, P, P, AR], optional=True)
                f[f"spgpu{s}hellspmvHalo"] = _sig(d, f"spgpu{s}hellspmvHalo", None,
                    [H, P, P, T, P, P, c_int, P, P, c_int, c_int, P, T, c_int, c_int, LK, U], optional=True)
                f[f"spgpu{s}hellspmvHaloDot"] = _sig(d, f"spgpu{s}hellspmvHaloDot", None,
                    [H, P, P, P, c_int, P, P, c_int, c_int, P, c_int, c_int, LK, U, P, AR], optional=True)
                f[f"spgpu{s}hdiaspmvHalo"] = _sig(d, f"spgpu{s}hdiaspmvHalo", None,
                    [H, P, P, T, P, P, c_int, P, c_int, c_int, P, T, c_int, LK, U], optional=True)
                f[f"spgpu{s}hdiaspmvHaloDot"] = _sig(d, f"spgpu{s}hdiaspmvHaloDot", None,
                    [H, P, P, P, c_int, P, c_int, c_int, P, c_int, LK, U, P, AR], optional=True)
            f["spgpuCsrToHellLayoutDevice"] = _sig(d, "spgpuCsrToHellLayoutDevice", c_int,
                [H, c_int, P, c_int, P, P, ctypes.POINTER(ctypes.c_longlong)], optional=True)
            for s in FLOAT_SYMS:
                f[f"spgpu{s}csrToHellDevice"] = _sig(d, f"spgpu{s}csrToHellDevice", None,
                    [H, c_int, P, P, P, c_int, c_int, P, c_int, P, P], optional=True)
            f["spgpuCsrToOhellLayoutDevice"] = _sig(d, "spgpuCsrToOhellLayoutDevice", c_int,
                [H, c_int, P, c_int, P, P, P, ctypes.POINTER(ctypes.c_longlong)], optional=True)
            f["spgpuHdiaHackOffsetsFromCooDevice"] = _sig(d, "spgpuHdiaHackOffsetsFromCooDevice", c_int,
                [H, ctypes.POINTER(c_int), P, c_int, c_int, c_int, c_int, P, P, c_int], optional=True)
            for s in FLOAT_SYMS:
                f[f"spgpu{s}csrToOhellDevice"] = _sig(d, f"spgpu{s}csrToOhellDevice", None,
                    [H, c_int, P, P, P, c_int, c_int, P, P, c_int, P, P], optional=True)
                f[f"spgpu{s}cooToHdiaDevice"] = _sig(d, f"spgpu{s}cooToHdiaDevice", c_int,
                    [H, P, P, P, c_int, c_int, c_int, c_int, P, P, P, c_int], optional=True)
            f["spgpuIpcGetHandle"] = _sig(d, "spgpuIpcGetHandle", c_int, [P, P], optional=True)
            f["spgpuIpcOpenHandle"] = _sig(d, "spgpuIpcOpenHandle", c_int,
                [P, ctypes.POINTER(c_void_p)], optional=True)
            f["spgpuIpcCloseHandle"] = _sig(d, "spgpuIpcCloseHandle", c_int, [P], optional=True)
            f["spgpuDeviceAlloc"] = _sig(d, "spgpuDeviceAlloc", c_int,
                [ctypes.POINTER(c_void_p), c_size_t], optional=True)
            f["spgpuDeviceFree"] = _sig(d, "spgpuDeviceFree", c_int, [P], optional=True)
            f["spgpuHaloPush"] = _sig(d, "spgpuHaloPush", None, [H, P, P, c_size_t, P, U], optional=True)
            f["spgpuDhaloPush"] = _sig(d, "spgpuDhaloPush", None, [H, P, P, c_int, P, U], optional=True)
            f["spgpuWaitFlag"] = _sig(d, "spgpuWaitFlag", None, [H, P, U], optional=True)
            f["spgpuHaloExchange"] = _sig(d, "spgpuHaloExchange", None,
                [H, P, P, P, P, c_size_t, P, P, P, P, P, P, U], optional=True)
            f["spgpuDhaloExchange"] = _sig(d, "spgpuDhaloExchange", None,
                [H, P, P, P, P, c_int, P, P, P, P, P, P, U], optional=True)
            f["spgpuHaloAck"] = _sig(d, "spgpuHaloAck", None, [H, P, P, U], optional=True)
            f["spgpuSetSeqCounters"] = _sig(d, "spgpuSetSeqCounters", c_int, [H, P, P], optional=True)
            f["spgpuHaloSeqAdvance"] = _sig(d, "spgpuHaloSeqAdvance", None, [H], optional=True)
            f["spgpuHaloTraceRead"] = _sig(d, "spgpuHaloTraceRead", c_int, [H, P, c_int, c_int], optional=True)

            # --- C-level multi-GPU API (include/spgpu_mg.h)
            PP = ctypes.POINTER(c_void_p)
            f["spgpuMgCreate"] = _sig(d, "spgpuMgCreate", c_int, [PP, ctypes.POINTER(c_int), c_int], optional=True)
            f["spgpuMgDestroy"] = _sig(d, "spgpuMgDestroy", None, [P], optional=True)
            f["spgpuMgWorld"] = _sig(d, "spgpuMgWorld", c_int, [P], optional=True)
            f["spgpuMgRankHandle"] = _sig(d, "spgpuMgRankHandle", c_void_p, [P, c_int], optional=True)
            f["spgpuMgSetExchange"] = _sig(d, "spgpuMgSetExchange", c_int, [P, c_int], optional=True)
            f["spgpuMgExchange"] = _sig(d, "spgpuMgExchange", c_int, [P], optional=True)
            f["spgpuMgSynchronize"] = _sig(d, "spgpuMgSynchronize", c_int, [P], optional=True)
            f["spgpuMgHellPlan"] = _sig(d, "spgpuMgHellPlan", c_int,
                [c_int, P, c_int, P, P, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)], optional=True)
            f["spgpuMgHdiaPlan"] = _sig(d, "spgpuMgHdiaPlan", c_int,
                [c_int, c_int, P, P, c_int, P, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)], optional=True)
            f["spgpuMgHellCreateFromBlocks"] = _sig(d, "spgpuMgHellCreateFromBlocks", c_int,
                [P, PP, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), PP, PP, PP, PP,
                 ctypes.POINTER(ctypes.c_longlong), c_int], optional=True)
            f["spgpuMgMatrixDestroy"] = _sig(d, "spgpuMgMatrixDestroy", None, [P], optional=True)
            f["spgpuMgMatrixHalo"] = _sig(d, "spgpuMgMatrixHalo", c_int, [P], optional=True)
            f["spgpuMgMatrixRows"] = _sig(d, "spgpuMgMatrixRows", c_int, [P], optional=True)
            f["spgpuMgMatrixRowBlock"] = _sig(d, "spgpuMgMatrixRowBlock", None,
                [P, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)], optional=True)
            f["spgpuMgVectorCreate"] = _sig(d, "spgpuMgVectorCreate", c_int, [P, PP], optional=True)
            f["spgpuMgVectorDestroy"] = _sig(d, "spgpuMgVectorDestroy", None, [P], optional=True)
            f["spgpuMgVectorSet"] = _sig(d, "spgpuMgVectorSet", c_int, [P, P], optional=True)
            f["spgpuMgVectorGet"] = _sig(d, "spgpuMgVectorGet", c_int, [P, P], optional=True)
            f["spgpuMgVectorLocal"] = _sig(d, "spgpuMgVectorLocal", c_void_p, [P, c_int], optional=True)
            f["spgpuMgDcgCreate"] = _sig(d, "spgpuMgDcgCreate", c_int, [P, PP], optional=True)
            f["spgpuMgDcgStart"] = _sig(d, "spgpuMgDcgStart", c_int, [P, P, ctypes.POINTER(c_double)], optional=True)
            f["spgpuMgDcgStep"] = _sig(d, "spgpuMgDcgStep", c_int, [P, c_int, ctypes.POINTER(c_double)], optional=True)
            f["spgpuMgDcgSolution"] = _sig(d, "spgpuMgDcgSolution", c_void_p, [P], optional=True)
            f["spgpuMgDcgDestroy"] = _sig(d, "spgpuMgDcgDestroy", None, [P], optional=True)
            for s in FLOAT_SYMS:
                t = TYPES[s]
                T, R = t.ctype, t.rtype
                f[f"spgpuMg{s}hellCreate"] = _sig(d, f"spgpuMg{s}hellCreate", c_int,
                    [P, PP, P, P, c_int, P, P, c_int, c_int, c_int, c_int], optional=True)
                f[f"spgpuMg{s}hdiaCreate"] = _sig(d, f"spgpuMg{s}hdiaCreate", c_int,
                    [P, PP, P, P, c_int, P, c_int, c_int], optional=True)
                for op in ("hellspmv", "hdiaspmv", "spmv"):
                    f[f"spgpuMg{s}{op}"] = _sig(d, f"spgpuMg{s}{op}", c_int, [P, P, P, T, P, P, T], optional=True)
                f[f"spgpuMg{s}dot"] = _sig(d, f"spgpuMg{s}dot", c_int, [P, ctypes.POINTER(T), P, P], optional=True)
                f[f"spgpuMg{s}nrm2"] = _sig(d, f"spgpuMg{s}nrm2", c_int, [P, ctypes.POINTER(R), P], optional=True)
                f[f"spgpuMg{s}axpby"] = _sig(d, f"spgpuMg{s}axpby", c_int, [P, P, T, P, T, P], optional=True)

    def __getattr__(self, name):
        try:
            fn = self.__dict__["fn"][name]
        except KeyError:
            raise AttributeError(name) from None
        if fn is None:
            raise AttributeError(f"{name} is not exported by {self.path}")
        return fn

    def has(self, name) -> bool:
        return hasattr(self.dll, name)


_default = None


def lib() -> SpgpuLib:
    """The product library (loaded once)."""
    global _default
    if _default is None:
        _default = SpgpuLib(LIB_PATH)
    return _default


def ptr(a) -> int:
    """Address of a numpy array / torch tensor / None (NULL)."""
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)
