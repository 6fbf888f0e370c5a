"""ctypes binding of the C ABI declared in include/parsmurf_b200.h (libparsmurf_b200.so, built in-tree by
parsmurf_b200/csrc/Makefile for sm_100a).  There is no CPU fallback: importing works anywhere (so the build and
symbol checks run on a CPU box) but every compute entry needs a CUDA device."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PSM_BUILD_DIR: an alternative in-tree build directory (kernel experiments built with `make OUT=... EXTRA=-D...`)
_BUILD = os.environ.get("PSM_BUILD_DIR") or os.path.join(_HERE, "_build")
LIB_PATH = os.path.join(_BUILD, "libparsmurf_b200.so")
HOST_LIB_PATH = os.path.join(_BUILD, "libparsmurf_host.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "parsmurf_b200.h")

KERNEL_CLASSES = ["knn", "assemble", "feature_sort", "bootstrap", "build_lists", "split_search", "split_select",
                  "level_finalize", "partition_lists", "predict", "curves", "gather"]

ERRORS = {-1: "PSM_E_CUDA", -2: "PSM_E_ARG", -3: "PSM_E_TOOFEW_POS", -4: "PSM_E_LOCALK", -5: "PSM_E_NEIGHBOUR",
          -6: "PSM_E_CHUNK", -7: "PSM_E_LIMIT", -8: "PSM_E_OVERFLOW"}

dp = C.POINTER(C.c_double)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)
u32, u64, vp = C.c_uint32, C.c_uint64, C.c_void_p


class PsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "PSM_E?"), code, msg))
        self.code = code


def build(verbose=False):
    """Compile the CUDA library and the C++ host library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    subprocess.check_call(args + ([] if verbose else ["-s"]))
    host_mk = os.path.join(_HERE, "host", "Makefile")
    if os.path.exists(host_mk):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "host"), "-j8"] + ([] if verbose else ["-s"]))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("parsmurf_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.psm_version.restype = C.c_char_p
        L.psm_last_error.restype = C.c_char_p
        L.psm_last_error.argtypes = [vp]
        L.psm_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.psm_ctx_destroy.argtypes = [vp]
        L.psm_set_stream.argtypes = [vp, vp]
        L.psm_set_batch_budget.argtypes = [vp, u64]
        L.psm_sync.argtypes = [vp]
        L.psm_dataset_upload.argtypes = [vp, dp, u32, u32]
        L.psm_dataset_attach_device.argtypes = [vp, vp, u32, u32]
        L.psm_fold_begin.argtypes = [vp, u32p, u32, u32p, u32, u32p, u32, u32p, u32]
        L.psm_folds_upload.argtypes = [vp, u32, u32p, u32p, u32p, u32p]
        L.psm_fold_begin_device.argtypes = [vp, u32, u32p]
        L.psm_fold_indices_get.argtypes = [vp, u32p, u32p, u32p, u32p, u32p]
        L.psm_fold_knn.argtypes = [vp, u32]
        L.psm_fold_knn_get.argtypes = [vp, i32p, dp]
        L.psm_fold_knn_path.argtypes = [vp, C.POINTER(C.c_int), u32p]
        L.psm_fold_knn_set.argtypes = [vp, i32p, u32]
        L.psm_batch_begin.argtypes = [vp, u32, u32, u32, u32, u32]
        L.psm_batch_assemble.argtypes = [vp, i32p, u64]
        L.psm_batch_assemble_glibc.argtypes = [vp, u32p]
        L.psm_glibc_jump.argtypes = [u32p, u64]
        L.psm_glibc_fill_device.argtypes = [vp, u32p, u64, u64, i32p, u64]
        L.psm_batch_capacity.argtypes = [vp, u32, u32, u32, u32, u32, u32, u32p]
        L.psm_batch_get_training.argtypes = [vp, u32, u32p, dp]
        L.psm_batch_train.argtypes = [vp, u32, u32, u32p, u64p, u32p, u32]
        L.psm_batch_tree_nodes.argtypes = [vp, u32, u32, u64p]
        L.psm_batch_export_tree.argtypes = [vp, u32, u32, u64p, u64p, u64p, dp, dp]
        L.psm_batch_get_inbag.argtypes = [vp, u32, u32, u32p]
        L.psm_batch_predict_accumulate.argtypes = [vp, u32]
        L.psm_fold_scores_get.argtypes = [vp, dp, dp]
        L.psm_fold_scores_devptr.argtypes = [vp, C.POINTER(vp), u32p]
        L.psm_fold_scatter_host.argtypes = [vp, dp, dp]
        L.psm_curves.argtypes = [vp, u32p, dp, u64, C.c_int, dp]
        L.psm_curves_info.argtypes = [vp, u64p, C.POINTER(C.c_int), dp]
        L.psm_fold_run_partitions.argtypes = [vp] + [u32] * 9 + [i32p]
        L.psm_profile_enable.argtypes = [vp, C.c_int]
        L.psm_profile_classes.argtypes = [vp, C.c_uint32]
        L.psm_profile_reset.argtypes = [vp]
        L.psm_profile_get.argtypes = [vp, dp, u64p]
        L.psm_stats_get.argtypes = [vp, u64p, u64p, u64p]
        _lib = L
    return _lib


def glibc_jump(window31, n):
    """advance a chronological 31-word glibc rand() window by n draws (host, O(log n))"""
    w = np.ascontiguousarray(window31, np.uint32).copy()
    lib().psm_glibc_jump(w.ctypes.data_as(u32p), n)
    return w


def glibc_initial_window(seed=1):
    """window of a freshly seeded glibc generator (srand(seed)): r[313..343] in the linear indexing of random_r.c"""
    r = [seed if seed else 1]
    for i in range(1, 31):
        r.append((16807 * r[i - 1]) % 2147483647)
    r += r[0:3]
    for i in range(34, 344):
        r.append((r[i - 31] + r[i - 3]) & 0xFFFFFFFF)
    return np.array(r[344 - 31:344], np.uint32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Context:
    """One GPU's pipeline context (psm_ctx).  Methods mirror the C entry points one to one."""

    def __init__(self, device=0, stream=None):
        self.L = lib()
        h = vp()
        rc = self.L.psm_ctx_create(device, C.byref(h))
        if rc != 0:
            raise PsmError(rc, "psm_ctx_create failed: no usable CUDA device %d (parsmurf_b200 has no CPU fallback)" % device)
        self.h = h
        if stream is not None:
            self.set_stream(stream)

    def _ck(self, rc):
        if rc != 0:
            raise PsmError(rc, self.L.psm_last_error(self.h).decode())

    def folds_upload(self, pos_list, pos_off, neg_list, neg_off):
        a = [np.ascontiguousarray(v, np.uint32) for v in (pos_list, pos_off, neg_list, neg_off)]
        self._ck(self.L.psm_folds_upload(self.h, len(a[1]) - 1, _p(a[0], u32p), _p(a[1], u32p), _p(a[2], u32p), _p(a[3], u32p)))

    def fold_begin_device(self, fold, shuffle_window31=None):
        w = None if shuffle_window31 is None else np.ascontiguousarray(shuffle_window31, np.uint32)
        self._ck(self.L.psm_fold_begin_device(self.h, fold, _p(w, u32p)))

    def fold_indices_get(self):
        s = np.zeros(4, np.uint32)
        self._ck(self.L.psm_fold_indices_get(self.h, _p(s, u32p), None, None, None, None))
        a = [np.zeros(int(v), np.uint32) for v in s]
        self._ck(self.L.psm_fold_indices_get(self.h, _p(s, u32p), *[_p(v, u32p) for v in a]))
        return dict(trngPos=a[0], trngNeg=a[1], testPos=a[2], testNeg=a[3])

    def close(self):
        if getattr(self, "h", None):
            self.L.psm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.psm_set_stream(self.h, vp(cuda_stream_ptr)))

    def set_batch_budget(self, nbytes):
        self._ck(self.L.psm_set_batch_budget(self.h, nbytes))

    def sync(self):
        self._ck(self.L.psm_sync(self.h))

    def dataset_upload(self, x):
        assert x.dtype == np.float64 and x.flags.c_contiguous and x.ndim == 2
        self._x_keep = x
        self.n, self.m = x.shape[0], x.shape[1] - 1
        self._ck(self.L.psm_dataset_upload(self.h, _p(x, dp), self.n, self.m))

    def dataset_attach_device(self, dev_ptr, n, m):
        self.n, self.m = n, m
        self._ck(self.L.psm_dataset_attach_device(self.h, vp(dev_ptr), n, m))

    def fold_begin(self, trngPos, trngNeg, testPos, testNeg):
        a = [np.ascontiguousarray(v, np.uint32) for v in (trngPos, trngNeg, testPos, testNeg)]
        self.P, self.Nneg, self.Nte = len(a[0]), len(a[1]), len(a[2]) + len(a[3])
        self._ck(self.L.psm_fold_begin(self.h, _p(a[0], u32p), len(a[0]), _p(a[1], u32p), len(a[1]), _p(a[2], u32p), len(a[2]),
                                       _p(a[3], u32p), len(a[3])))

    def fold_knn(self, k):
        self._ck(self.L.psm_fold_knn(self.h, k))
        self.localk = min(k, self.P)

    def fold_knn_path(self):
        """(path, incomplete): 0 exact kernel, 1 tensor-core candidates + exact rerank proven complete, 2 tensor path fell back"""
        path = C.c_int(0); bad = np.zeros(1, np.uint32)
        self._ck(self.L.psm_fold_knn_path(self.h, C.byref(path), _p(bad, u32p)))
        return int(path.value), int(bad[0])

    def fold_knn_get(self):
        idx = np.zeros((self.P, self.localk), np.int32); dist = np.zeros((self.P, self.localk))
        self._ck(self.L.psm_fold_knn_get(self.h, _p(idx, i32p), _p(dist, dp)))
        return idx, dist

    def fold_knn_set(self, nn):
        nn = np.ascontiguousarray(nn, np.int32)
        self.localk = nn.shape[1]
        self._ck(self.L.psm_fold_knn_set(self.h, _p(nn, i32p), nn.shape[1]))

    def batch_begin(self, nParts, fp, ratio, p0, p1):
        self._ck(self.L.psm_batch_begin(self.h, nParts, fp, ratio, p0, p1))

    def batch_assemble(self, draws):
        draws = None if draws is None else np.ascontiguousarray(draws, np.int32)
        self._ck(self.L.psm_batch_assemble(self.h, _p(draws, i32p), 0 if draws is None else draws.size))

    def batch_assemble_glibc(self, window31):
        w = np.ascontiguousarray(window31, np.uint32)
        assert w.size == 31
        self._ck(self.L.psm_batch_assemble_glibc(self.h, _p(w, u32p)))

    def glibc_fill_device(self, window31, count, skip=0, n_out=None):
        """outputs [skip, skip + n_out) of the next `count` rand() draws, generated on the device"""
        w = np.ascontiguousarray(window31, np.uint32)
        assert w.size == 31
        n_out = count - skip if n_out is None else n_out
        out = np.empty(n_out, np.int32)
        self._ck(self.L.psm_glibc_fill_device(self.h, _p(w, u32p), count, skip, _p(out, i32p), n_out))
        return out

    def batch_capacity(self, nParts, fp, ratio, nTrees, mtry, p0):
        c = u32()
        self._ck(self.L.psm_batch_capacity(self.h, nParts, fp, ratio, nTrees, mtry, p0, C.byref(c)))
        return c.value

    def batch_get_training(self, p):
        sizes = np.zeros(3, np.uint32)
        self._ck(self.L.psm_batch_get_training(self.h, p, _p(sizes, u32p), None))
        out = np.zeros((int(sizes[2]), self.m + 1))
        self._ck(self.L.psm_batch_get_training(self.h, p, _p(sizes, u32p), _p(out, dp)))
        return out, dict(trngPos=int(sizes[0]), trngNeg=int(sizes[1]), Ntr=int(sizes[2]))

    def batch_train(self, nTrees, mtry, forest_seeds, inj_bootstrap=None, inj_cand=None):
        seeds = None if forest_seeds is None else np.ascontiguousarray(forest_seeds, np.uint32)
        ib = None if inj_bootstrap is None else np.ascontiguousarray(inj_bootstrap, np.uint64)
        ic = None if inj_cand is None else np.ascontiguousarray(inj_cand, np.uint32)
        nodes = 0 if ic is None else ic.shape[1]
        self._ck(self.L.psm_batch_train(self.h, nTrees, mtry, _p(seeds, u32p), _p(ib, u64p), _p(ic, u32p), nodes))

    def batch_export_tree(self, p, t):
        nn = u64()
        self._ck(self.L.psm_batch_tree_nodes(self.h, p, t, C.byref(nn)))
        nn = nn.value
        left = np.zeros(nn, np.uint64); right = np.zeros(nn, np.uint64); var = np.zeros(nn, np.uint64)
        val = np.zeros(nn); frac = np.zeros((nn, 2))
        self._ck(self.L.psm_batch_export_tree(self.h, p, t, _p(left, u64p), _p(right, u64p), _p(var, u64p), _p(val, dp), _p(frac, dp)))
        return dict(left=left, right=right, var=var, val=val, frac=frac)

    def batch_get_inbag(self, p, t, Ntr):
        c = np.zeros(Ntr, np.uint32)
        self._ck(self.L.psm_batch_get_inbag(self.h, p, t, _p(c, u32p)))
        return c

    def batch_predict_accumulate(self, nPartsTotal):
        self._ck(self.L.psm_batch_predict_accumulate(self.h, nPartsTotal))

    def fold_scores_get(self):
        c1 = np.zeros(self.Nte); c2 = np.zeros(self.Nte)
        self._ck(self.L.psm_fold_scores_get(self.h, _p(c1, dp), _p(c2, dp)))
        return c1, c2

    def fold_scores_devptr(self):
        ptr = vp(); nte = u32()
        self._ck(self.L.psm_fold_scores_devptr(self.h, C.byref(ptr), C.byref(nte)))
        return ptr.value, nte.value

    def fold_scatter_host(self, class1, class2):
        self._ck(self.L.psm_fold_scatter_host(self.h, _p(class1, dp), _p(class2, dp)))

    def curves(self, labels, scores, tie_mode=0):
        labels = np.ascontiguousarray(labels, np.uint32); scores = np.ascontiguousarray(scores, np.float64)
        out = np.zeros(2)
        self._ck(self.L.psm_curves(self.h, _p(labels, u32p), _p(scores, dp), len(labels), tie_mode, _p(out, dp)))
        return float(out[0]), float(out[1])

    def curves_info(self):
        """(mixed_tie_pairs, used_host_sort, (auroc spread, auprc spread)) of the last curves call (tie_mode 2)"""
        a = u64(); b = C.c_int(0); sp = np.zeros(2)
        self._ck(self.L.psm_curves_info(self.h, C.byref(a), C.byref(b), _p(sp, dp)))
        return int(a.value), int(b.value), (float(sp[0]), float(sp[1]))

    def fold_run_partitions(self, nParts, fp, ratio, nTrees, mtry, seed, fold, p0, p1, draws):
        draws = None if draws is None else np.ascontiguousarray(draws, np.int32)
        self._ck(self.L.psm_fold_run_partitions(self.h, nParts, fp, ratio, nTrees, mtry, seed, fold, p0, p1, _p(draws, i32p)))

    def profile_enable(self, on=True):
        self._ck(self.L.psm_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._ck(self.L.psm_profile_reset(self.h))

    def profile_get(self):
        ms = np.zeros(len(KERNEL_CLASSES)); ln = np.zeros(len(KERNEL_CLASSES), np.uint64)
        self._ck(self.L.psm_profile_get(self.h, _p(ms, dp), _p(ln, u64p)))
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(KERNEL_CLASSES)}

    def stats_get(self):
        a, b, c = u64(), u64(), u64()
        self._ck(self.L.psm_stats_get(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(split_bytes=a.value, nodes=b.value, levels=c.value)
