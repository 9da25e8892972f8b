"""ctypes binding of libmmz.so (include/mmz.h) over PyTorch device memory.

PyTorch is used for device allocations, streams and (optionally)
torch.distributed; every computation of the step path happens inside the CUDA
library. There is NO CPU fallback: if the library is missing or CUDA is not
available this module raises — the hot path must never silently run elsewhere.
"""

import ctypes
import os
from typing import Optional, Tuple

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# MMZ_LIB selects another build of the same library (e.g. the libmmz_dbg.so of tools/build_debug.py)
LIB_PATH = os.path.join(_PKG_ROOT, os.environ.get("MMZ_LIB", "libmmz.so"))

MMZ_DONE, MMZ_TRUNCATED, MMZ_UNSTABLE = 1, 2, 4
MMZ_AUTO_RESET = 1
LAYOUT_ENV_MAJOR, LAYOUT_SOA = 0, 1
ABI_VERSION = 7

_lib = None


class MmzError(RuntimeError):
    pass


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen libmmz.so and declare the signatures of include/mmz.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise MmzError(
            f"{path} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
            "There is no CPU fallback for the step path."
        )
    lib = ctypes.CDLL(path)
    vp, i32, u32, u64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t
    ip = ctypes.POINTER(ctypes.c_int)
    sig = {
        "mmz_create": ([vp, sz, i32, i32, u32, ctypes.POINTER(vp)], i32),
        "mmz_dims": ([vp, ip, ip, ip, ip, ip], i32),
        "mmz_kernel_config": ([vp, ip, ip, ip, ip, ip], i32),
        "mmz_kernel_name": ([vp], ctypes.c_char_p),
        "mmz_set_env_offset": ([vp, i32], i32),
        "mmz_set_obs_peers": ([vp, ctypes.POINTER(vp), i32, ctypes.c_int64, i32], i32),
        "mmz_set_step_diag": ([vp, vp], i32),
        "mmz_reset": ([vp, vp, u64, vp, vp], i32),
        "mmz_step": ([vp, vp, vp, vp, vp, vp, vp], i32),
        "mmz_step_k": ([vp, i32, vp, vp, vp, vp, vp, vp], i32),
        "mmz_step_host": ([vp, vp, vp, vp, vp, vp, vp], i32),
        "mmz_observe": ([vp, vp, vp], i32),
        "mmz_get_state": ([vp, i32, vp, vp, vp, vp], i32),
        "mmz_set_state": ([vp, i32, vp, vp, vp, vp], i32),
        "mmz_forward": ([vp, vp, vp, vp, vp], i32),
        "mmz_render": ([vp, i32, i32, i32, i32, vp, vp], i32),
        "mmz_launch_count": ([vp], u64),
        "mmz_last_error": ([], ctypes.c_char_p),
        "mmz_destroy": ([vp], None),
        "mmz_abi_version": ([], i32),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.argtypes, fn.restype = argtypes, restype
    if lib.mmz_abi_version() != ABI_VERSION:
        raise MmzError(f"libmmz ABI {lib.mmz_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    if path == LIB_PATH:
        _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "mmz_create", "mmz_dims", "mmz_kernel_config", "mmz_kernel_name", "mmz_set_env_offset", "mmz_set_obs_peers", "mmz_set_step_diag", "mmz_reset", "mmz_step", "mmz_step_k", "mmz_step_host", "mmz_observe", "mmz_get_state",
    "mmz_set_state", "mmz_forward", "mmz_render", "mmz_launch_count", "mmz_last_error", "mmz_destroy", "mmz_abi_version",
)


def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


class BatchedSim:
    """N lock-step environments on one CUDA device (one `mmz_handle`)."""

    def __init__(self, model, num_envs: int, device="cuda:0", auto_reset: bool = False, env_offset: int = 0):
        import torch

        self.torch = torch
        if not torch.cuda.is_available():
            raise MmzError("CUDA is not available: the maze step path is GPU-only (no CPU fallback)")
        self.lib = load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise MmzError(f"device must be a CUDA device, got {device}")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.model = model
        self.n = int(num_envs)
        blob = model.blob(4)
        self._h = ctypes.c_void_p()
        flags = MMZ_AUTO_RESET if auto_reset else 0
        self._check(self.lib.mmz_create(blob, len(blob), self.n, self.index, flags, ctypes.byref(self._h)))
        dims = [ctypes.c_int() for _ in range(5)]
        self._check(self.lib.mmz_dims(self._h, *[ctypes.byref(d) for d in dims]))
        _, self.nq, self.nv, self.nu, self.obs_dim = (d.value for d in dims)
        cfg = [ctypes.c_int() for _ in range(5)]
        self._check(self.lib.mmz_kernel_config(self._h, *[ctypes.byref(c) for c in cfg]))
        self.kernel_config = dict(zip(
            ("lanes_per_env", "threads_per_block", "smem_bytes", "envs_per_sm", "floats_per_env"),
            (c.value for c in cfg)))
        self.kernel_config["kernel"] = self.lib.mmz_kernel_name(self._h).decode()
        if env_offset:
            self._check(self.lib.mmz_set_env_offset(self._h, int(env_offset)))
        dev = self.device
        self.obs = torch.empty((self.n, self.obs_dim), dtype=torch.float32, device=dev)
        self.reward = torch.empty((self.n,), dtype=torch.float32, device=dev)
        self.done = torch.empty((self.n,), dtype=torch.uint8, device=dev)
        self.info = torch.empty((self.n, 4), dtype=torch.float32, device=dev)

    # -- helpers -------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise MmzError(f"libmmz error {rc}: {self.lib.mmz_last_error().decode()}")

    def _stream(self) -> int:
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _f32(self, x, shape) -> "torch.Tensor":
        t = self.torch.as_tensor(x, dtype=self.torch.float32, device=self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    # -- API -----------------------------------------------------------------
    def reset(self, seed: int = 0, mask=None):
        m = None
        if mask is not None:
            m = self.torch.as_tensor(mask, device=self.device).to(self.torch.uint8).contiguous()
        self._check(self.lib.mmz_reset(self._h, _ptr(m), int(seed), self.obs.data_ptr(), self._stream()))
        return self.obs

    def step(self, action):
        a = self._f32(action, (self.n, self.nu))
        self._check(self.lib.mmz_step(self._h, a.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(),
                                      self.done.data_ptr(), self.info.data_ptr(), self._stream()))
        return self.obs, self.reward, self.done, self.info

    def _buffers(self, action, obs, reward, done, info, host: bool):
        """The C ABI takes raw addresses: a wrong dtype, shape, stride or device would be a silent out-of-bounds access."""
        t = self.torch
        # a rollout loop passes the same tensors again and again: a set that was validated is recognised by identity,
        # address and size (is_pinned() alone costs more than the launch of a small step)
        key = (host,) + tuple((id(x), x.data_ptr(), x.dtype, x.shape, x.is_contiguous()) if t.is_tensor(x) else None
                              for x in (action, obs, reward, done, info))
        if key == getattr(self, "_validated", None):
            return
        self._validated = None
        want = (("action", action, (self.n, self.nu), t.float32), ("obs", obs, (self.n, self.obs_dim), t.float32),
                ("reward", reward, (self.n,), t.float32), ("done", done, (self.n,), t.uint8), ("info", info, (self.n, 4), t.float32))
        for name, x, shape, dtype in want:
            if x is None and name == "info":
                continue
            if not t.is_tensor(x) or x.dtype != dtype or tuple(x.shape) != shape or not x.is_contiguous():
                raise ValueError(f"{name}: expected a contiguous {dtype} tensor of shape {shape}, got "
                                 f"{type(x).__name__} {getattr(x, 'dtype', None)} {tuple(getattr(x, 'shape', ()))}")
            if host:
                if x.device.type != "cpu" or not x.is_pinned():
                    raise ValueError(f"{name}: step_host needs pinned host memory (tensor.pin_memory()), got {x.device}")
            elif x.device != self.device and not (x.device.type == "cuda" and x.device.index == self.index):
                raise ValueError(f"{name}: tensor on {x.device}, the environments live on {self.device}")
        self._validated = key

    def step_into(self, action, obs, reward, done, info=None):
        """Step with caller-provided output tensors (bench / multi-buffering)."""
        self._buffers(action, obs, reward, done, info, host=False)
        self._check(self.lib.mmz_step(self._h, action.data_ptr(), obs.data_ptr(), reward.data_ptr(),
                                      done.data_ptr(), _ptr(info), self._stream()))

    def step_k(self, actions, out=None):
        """K steps in one host call (include/mmz.h: mmz_step_k): `actions` [K, N, nu] on the device. Returns
        (obs [K, N, obs_dim], reward [K, N], done [K, N] uint8, info [K, N, 4]); pass the previous result as `out` to reuse
        its buffers (and the recorded CUDA graph)."""
        t = self.torch
        if not t.is_tensor(actions) or actions.dim() != 3 or tuple(actions.shape[1:]) != (self.n, self.nu):
            raise ValueError(f"actions: expected a [K, {self.n}, {self.nu}] tensor")
        a = actions.to(device=self.device, dtype=t.float32).contiguous()
        K = int(a.shape[0])
        if out is None or out[0].shape[0] != K:
            out = (t.empty((K, self.n, self.obs_dim), dtype=t.float32, device=self.device),
                   t.empty((K, self.n), dtype=t.float32, device=self.device),
                   t.empty((K, self.n), dtype=t.uint8, device=self.device),
                   t.empty((K, self.n, 4), dtype=t.float32, device=self.device))
        self._k_actions = a  # the graph holds its address
        self._check(self.lib.mmz_step_k(self._h, K, a.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                        out[3].data_ptr(), self._stream()))
        return out

    def step_host(self, h_action, h_obs, h_reward, h_done, h_info=None):
        """End-to-end step through pinned host tensors (synchronises the stream)."""
        self._buffers(h_action, h_obs, h_reward, h_done, h_info, host=True)
        self._check(self.lib.mmz_step_host(self._h, h_action.data_ptr(), h_obs.data_ptr(), h_reward.data_ptr(),
                                           h_done.data_ptr(), _ptr(h_info), self._stream()))

    def observe(self):
        self._check(self.lib.mmz_observe(self._h, self.obs.data_ptr(), self._stream()))
        return self.obs

    def get_state(self, layout: int = LAYOUT_ENV_MAJOR):
        t = self.torch
        shape_q = (self.n, self.nq) if layout == LAYOUT_ENV_MAJOR else (self.nq, self.n)
        shape_v = (self.n, self.nv) if layout == LAYOUT_ENV_MAJOR else (self.nv, self.n)
        qpos = t.empty(shape_q, dtype=t.float32, device=self.device)
        qvel = t.empty(shape_v, dtype=t.float32, device=self.device)
        step = t.empty((self.n,), dtype=t.int32, device=self.device)
        self._check(self.lib.mmz_get_state(self._h, layout, qpos.data_ptr(), qvel.data_ptr(), step.data_ptr(),
                                           self._stream()))
        return qpos, qvel, step

    def set_state(self, qpos=None, qvel=None, t=None, layout: int = LAYOUT_ENV_MAJOR):
        shape_q = (self.n, self.nq) if layout == LAYOUT_ENV_MAJOR else (self.nq, self.n)
        shape_v = (self.n, self.nv) if layout == LAYOUT_ENV_MAJOR else (self.nv, self.n)
        q = None if qpos is None else self._f32(qpos, shape_q)
        v = None if qvel is None else self._f32(qvel, shape_v)
        s = None if t is None else self.torch.as_tensor(t, dtype=self.torch.int32, device=self.device).contiguous()
        self._check(self.lib.mmz_set_state(self._h, layout, _ptr(q), _ptr(v), _ptr(s), self._stream()))

    def forward(self, action) -> Tuple["torch.Tensor", "torch.Tensor"]:
        t = self.torch
        a = self._f32(action, (self.n, self.nu))
        qacc = t.empty((self.n, self.nv), dtype=t.float32, device=self.device)
        diag = t.empty((self.n, 4), dtype=t.int32, device=self.device)
        self._check(self.lib.mmz_forward(self._h, a.data_ptr(), qacc.data_ptr(), diag.data_ptr(), self._stream()))
        return qacc, diag

    def render(self, width: int = 256, height: int = 256, first_env: int = 0, count: Optional[int] = None):
        """Top-down RGB images of environments [first_env, first_env + count): uint8 [count, height, width, 3]."""
        t = self.torch
        count = self.n - first_env if count is None else int(count)
        rgb = t.empty((count, height, width, 3), dtype=t.uint8, device=self.device)
        self._check(self.lib.mmz_render(self._h, int(first_env), count, int(width), int(height), rgb.data_ptr(),
                                        self._stream()))
        return rgb

    def set_obs_peers(self, ptrs, row_offset: int, multicast: bool = False):
        """Fused observation gather (include/mmz.h: mmz_set_obs_peers): device addresses of the [total, obs_dim] float32
        gathered tensors of every rank (or one multicast address); `ptrs` empty turns it off."""
        ptrs = [int(p) for p in ptrs]
        arr = (ctypes.c_void_p * max(1, len(ptrs)))(*ptrs)
        self._check(self.lib.mmz_set_obs_peers(self._h, arr, len(ptrs), int(row_offset), int(bool(multicast))))

    def enable_step_diag(self, on: bool = True):
        """Per-env solver diagnostics of every following step: [N, 4] int32 (see include/mmz.h)."""
        t = self.torch
        self.step_diag = t.zeros((self.n, 4), dtype=t.int32, device=self.device) if on else None
        self._check(self.lib.mmz_set_step_diag(self._h, _ptr(self.step_diag)))
        return self.step_diag

    @property
    def launch_count(self) -> int:
        return int(self.lib.mmz_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mmz_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
