"""ctypes binding of oracle/cpu/liboracle_cpu.so (TEST INFRASTRUCTURE -- see oracle/__init__.py)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "cpu", "liboracle_cpu.so")
SRC = os.path.join(HERE, "cpu", "oracle_cpu.cpp")
_lib = None


def _host_signature():
    """the .so is built with -march=native: rebuild it when the host CPU differs from the build host
    (the build container and the GPU box are different machines)."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def load():
    global _lib
    if _lib is None:
        sig_file = SO + ".host"
        sig = _host_signature()
        stale = (not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC) or not os.path.exists(sig_file)
                 or open(sig_file).read().strip() != sig)
        if stale:
            subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-shared", "-fPIC", "-pthread",
                                   "-o", SO, SRC])
            with open(sig_file, "w") as f:
                f.write(sig)
        _lib = ctypes.CDLL(SO)
        _lib.orc_hardware_threads.restype = ctypes.c_int
    return _lib


def threads():
    return max(1, load().orc_hardware_threads())


def _sz(n):
    return ctypes.c_size_t(n)


def fp_bytes(cid):
    return 32 if cid == 1 else 48


def pairing_batch(cid, n, g1a, g2a, g1b=None, g2b=None, fexp=False, unity=False, nthreads=0):
    np_ = 2 if g1b is not None else 1
    osz = n if unity else n * 12 * fp_bytes(cid)
    out = ctypes.create_string_buffer(max(1, osz))
    rc = load().orc_pairing_batch(cid, np_, _sz(n), g1a, g2a, g1b, g2b, out, int(fexp), int(unity), nthreads or threads())
    if rc:
        raise RuntimeError("oracle pairing rc=%d" % rc)
    return out.raw[:osz]


def fexp_batch(cid, n, gt, nthreads=0):
    out = ctypes.create_string_buffer(max(1, n * 12 * fp_bytes(cid)))
    rc = load().orc_fexp_batch(cid, _sz(n), gt, out, nthreads or threads())
    if rc:
        raise RuntimeError("oracle fexp rc=%d" % rc)
    return out.raw[:n * 12 * fp_bytes(cid)]


def g1_mul_batch(cid, n, pts, ks, nthreads=0):
    out = ctypes.create_string_buffer(max(1, n * 2 * fp_bytes(cid)))
    rc = load().orc_g1_mul_batch(cid, _sz(n), pts, ks, out, nthreads or threads())
    if rc:
        raise RuntimeError("oracle g1_mul rc=%d" % rc)
    return out.raw[:n * 2 * fp_bytes(cid)]


def g1_mul2_batch(cid, n, P, e, Q, f, nthreads=0):
    out = ctypes.create_string_buffer(max(1, n * 2 * fp_bytes(cid)))
    rc = load().orc_g1_mul2_batch(cid, _sz(n), P, e, Q, f, out, nthreads or threads())
    if rc:
        raise RuntimeError("oracle g1_mul2 rc=%d" % rc)
    return out.raw[:n * 2 * fp_bytes(cid)]


def g1_msm(cid, n, pts, ks, nthreads=0):
    out = ctypes.create_string_buffer(2 * fp_bytes(cid))
    rc = load().orc_g1_msm(cid, _sz(n), pts, ks, out, nthreads or threads())
    if rc:
        raise RuntimeError("oracle msm rc=%d" % rc)
    return out.raw
