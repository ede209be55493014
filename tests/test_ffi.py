"""The XLA-FFI adapter (include/pupper_ffi.h, csrc/pupper_ffi.cc) driven through a hand-built call frame.

jaxlib is not installable here, so XLA itself never calls the handlers; these tests play XLA's part: they lay out the
XLA_FFI_CallFrame / XLA_FFI_Buffer structs of csrc/xla_ffi_stub.h with ctypes, provide the two API callbacks the handlers
use (stream lookup, error creation) and check (CPU) the blob arithmetic and the argument validation, (GPU) that a reset + steps
through the handlers -- DR blob, episode blob and per-device model registry included -- reproduce EnvRuntime bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

import common
from pupperv3_mjx_b200 import abi, domain_randomization as dr, prng, runtime

F32, U32 = 11, 8  # XLA_FFI_DataType


class Buffer(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("dtype", C.c_int), ("data", C.c_void_p),
                ("rank", C.c_int64), ("dims", C.POINTER(C.c_int64))]


class Args(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("size", C.c_int64), ("types", C.POINTER(C.c_int32)),
                ("args", C.POINTER(C.c_void_p))]


class Attrs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("size", C.c_int64), ("types", C.c_void_p),
                ("names", C.c_void_p), ("attrs", C.c_void_p)]


class ApiVersion(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("major", C.c_int), ("minor", C.c_int)]


class ErrorCreateArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("message", C.c_char_p), ("errc", C.c_int)]


class StreamGetArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("ctx", C.c_void_p), ("stream", C.c_void_p)]


ERR_CREATE = C.CFUNCTYPE(C.c_void_p, C.POINTER(ErrorCreateArgs))
STREAM_GET = C.CFUNCTYPE(C.c_void_p, C.POINTER(StreamGetArgs))


class Api(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("api_version", ApiVersion), ("internal_api", C.c_void_p),
                ("error_create", ERR_CREATE), ("error_get_message", C.c_void_p), ("error_destroy", C.c_void_p),
                ("handler_register", C.c_void_p), ("stream_get", STREAM_GET)]


class CallFrame(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("api", C.POINTER(Api)), ("ctx", C.c_void_p),
                ("stage", C.c_int), ("args", Args), ("rets", Args), ("attrs", Attrs), ("future", C.c_void_p)]


class FakeXla:
    """Owns the callbacks and keeps every ctypes object of a call alive."""

    def __init__(self, stream=0):
        self.errors, self.stream, self.keep = [], stream, []
        self._ec = ERR_CREATE(self._error_create)
        self._sg = STREAM_GET(self._stream_get)
        self.api = Api(struct_size=C.sizeof(Api), error_create=self._ec, stream_get=self._sg)

    def _error_create(self, a):
        self.errors.append((a.contents.message.decode(), a.contents.errc))
        return 0xDEAD  # any non-null token stands for the XLA_FFI_Error*

    def _stream_get(self, a):
        a.contents.stream = self.stream
        return None

    def buf(self, ptr, dims, dtype=F32):
        d = (C.c_int64 * max(len(dims), 1))(*dims)
        b = Buffer(struct_size=C.sizeof(Buffer), dtype=dtype, data=ptr, rank=len(dims), dims=d)
        self.keep += [d, b]
        return b

    def frame(self, args, rets, stage=3):
        def pack(bufs):
            arr = (C.c_void_p * max(len(bufs), 1))(*[C.addressof(b) for b in bufs])
            types = (C.c_int32 * max(len(bufs), 1))(*([1] * len(bufs)))
            self.keep += [arr, types]
            return Args(struct_size=C.sizeof(Args), size=len(bufs), types=types, args=arr)
        f = CallFrame(struct_size=C.sizeof(CallFrame), api=C.pointer(self.api), stage=stage, args=pack(args), rets=pack(rets))
        self.keep.append(f)
        return f


def _lib():
    lib = runtime.load_library()
    lib.PupperStepFfi.restype = C.c_void_p
    lib.PupperResetFfi.restype = C.c_void_p
    lib.PupperStepFfi.argtypes = [C.c_void_p]
    lib.PupperResetFfi.argtypes = [C.c_void_p]
    for f in ("PupperPolicyFfi", "PupperRolloutFfi"):
        getattr(lib, f).restype = C.c_void_p
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.pupper_ffi_register_policy.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int]
    for f in ("pupper_state_blob_words", "pupper_dr_blob_words", "pupper_episode_blob_words", "pupper_rand_blob_words"):
        getattr(lib, f).restype = C.c_int64
    return lib


def test_blob_layouts_match_the_runtime_state():
    lib = _lib()
    env = common.make_env()
    cfg = env.env_cfg
    n = 37
    stride = lib.pupper_blob_stride(n)
    assert stride == 64
    rows = (C.c_int32 * 14)()
    lib.pupper_state_rows(C.byref(cfg), rows)
    words = lib.pupper_state_blob_words(C.byref(cfg), n)
    assert words == sum(rows) * stride + n * cfg.observation_history * 36
    assert lib.pupper_dr_blob_words(n) == 58 * stride and lib.pupper_rand_blob_words(n) == 44 * stride
    assert lib.pupper_episode_blob_words(C.byref(cfg), n) == 79 * stride + n * cfg.observation_history * 36 + 24
    blob = np.zeros(words, np.float32)
    st = abi.PupperState()
    assert lib.pupper_state_blob_bind(C.byref(cfg), n, blob.ctypes.data_as(C.c_void_p), C.byref(st)) == 0
    off, base = 0, blob.ctypes.data
    for name, r in zip(abi.STATE_FIELDS, rows):
        assert getattr(st, name) == base + 4 * off, name
        off += r * stride
    assert st.obs == base + 4 * off and st.stride == stride
    # DR pack: env-major leaves (what domain_randomize returns) -> [58][stride]
    sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n))
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    fr, kp, kd = f(sys_v.geom_friction[:, 0, 0]), f(sys_v.actuator_gainprm[:, 0, 0]), f(-sys_v.actuator_biasprm[:, 0, 2])
    ipos, inertia, mass = f(sys_v.body_ipos[:, 1]), f(sys_v.body_inertia[:, 1:]), f(sys_v.body_mass[:, 1:])
    drb = np.full(58 * stride, np.nan, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.pupper_dr_blob_pack(n, p(fr), p(kp), p(kd), p(ipos), p(inertia), p(mass), p(drb)) == 0
    drb = drb.reshape(58, stride)
    np.testing.assert_array_equal(drb[0, :n], fr); np.testing.assert_array_equal(drb[2, :n], kd)
    np.testing.assert_array_equal(drb[3:6, :n], ipos.T); np.testing.assert_array_equal(drb[6:45, :n], inertia.reshape(n, 39).T)
    np.testing.assert_array_equal(drb[45:58, :n], mass.T)
    assert np.all(drb[:, n:] == 0)
    d = abi.PupperDR()
    assert lib.pupper_dr_blob_bind(n, p(drb), C.byref(d)) == 0
    assert d.body_mass == drb.ctypes.data + 4 * 45 * stride and d.stride == stride


def test_handlers_validate_the_call_frame_without_a_gpu():
    lib = _lib()
    x = FakeXla()
    one = np.zeros(24, np.float32)
    b = x.buf(one.ctypes.data, (2, 12))
    assert lib.PupperStepFfi(C.addressof(x.frame([b, b], [b]))) == 0xDEAD            # wrong operand count
    assert "5 arguments" in x.errors[-1][0]
    bad = x.buf(one.ctypes.data, (4, 6))
    assert lib.PupperStepFfi(C.addressof(x.frame([bad] * 5, [b] * 5))) == 0xDEAD     # action is not [n, 12]
    assert "f32[n, 12]" in x.errors[-1][0]
    k = x.buf(one.ctypes.data, (4, 2), F32)
    assert lib.PupperResetFfi(C.addressof(x.frame([k] * 3, [b] * 5))) == 0xDEAD      # keys must be u32
    assert "u32[n, 2]" in x.errors[-1][0]
    assert lib.PupperStepFfi(C.addressof(x.frame([b] * 5, [b] * 5, stage=1))) is None  # not the EXECUTE stage: no-op
    assert lib.pupper_ffi_register_model(0, None, None) == -1 and lib.pupper_ffi_unregister_model(99) == -1
    # policy / rollout handlers
    assert lib.PupperPolicyFfi(C.addressof(x.frame([b, b], [b]))) == 0xDEAD and "1 argument" in x.errors[-1][0]
    assert lib.PupperPolicyFfi(C.addressof(x.frame([x.buf(one.ctypes.data, (4, 6))], [x.buf(one.ctypes.data, (2, 12))]))) == 0xDEAD  # row counts differ
    assert "f32[n, in]" in x.errors[-1][0]
    assert lib.PupperRolloutFfi(C.addressof(x.frame([b] * 2, [b] * 7))) == 0xDEAD and "3 arguments and 7 results" in x.errors[-1][0]
    assert lib.PupperRolloutFfi(C.addressof(x.frame([b] * 3, [b] * 7))) == 0xDEAD and "f32[T, n, 12]" in x.errors[-1][0]  # action result is rank 2
    assert lib.PupperRolloutFfi(C.addressof(x.frame([b] * 3, [b] * 7, stage=1))) is None
    assert lib.pupper_ffi_register_policy(0, None, 72, 12) == -1 and lib.pupper_ffi_unregister_policy(99) == -1


@pytest.mark.gpu
def test_reset_and_steps_through_the_ffi_handlers_match_the_runtime():
    torch = pytest.importorskip("torch")
    lib = _lib()
    env = common.make_env()
    env.set_episode_params(50, 1)
    cfg = env.env_cfg
    n = 200
    dev = torch.device("cuda", 0)
    keys = common.env_keys(n)
    sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n))
    rt = runtime.EnvRuntime(env.model_desc, cfg, n, episode=True)     # the ctypes / torch path, as the reference
    rt.set_dr(sys_v)
    rt.reset(torch.from_numpy(keys.view(np.int32)).cuda())
    # the FFI path: everything lives in flat float32 blobs, as JAX would hold it
    assert lib.pupper_ffi_register_model(0, rt._model, C.byref(cfg)) == 0
    stride = lib.pupper_blob_stride(n)
    zeros = lambda w: torch.zeros(int(w), dtype=torch.float32, device=dev)
    state = zeros(lib.pupper_state_blob_words(C.byref(cfg), n))
    ep = zeros(lib.pupper_episode_blob_words(C.byref(cfg), n))
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    host_dr = np.zeros(58 * stride, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.pupper_dr_blob_pack(n, p(f(sys_v.geom_friction[:, 0, 0])), p(f(sys_v.actuator_gainprm[:, 0, 0])), p(f(-sys_v.actuator_biasprm[:, 0, 2])),
                            p(f(sys_v.body_ipos[:, 1])), p(f(sys_v.body_inertia[:, 1:])), p(f(sys_v.body_mass[:, 1:])), p(host_dr))
    drb = torch.from_numpy(host_dr).to(dev)
    reward, done, metrics = zeros(n), zeros(n), zeros(n * 19)
    d_keys = torch.from_numpy(keys.view(np.int32)).to(dev)
    empty = zeros(0)
    x = FakeXla(stream=torch.cuda.current_stream().cuda_stream)
    B = lambda t, dims, dt=F32: x.buf(t.data_ptr() if t.numel() else None, dims, dt)
    fr = x.frame([B(d_keys, (n, 2), U32), B(drb, (drb.numel(),)), B(empty, (0,))],
                 [B(state, (state.numel(),)), B(reward, (n,)), B(done, (n,)), B(metrics, (n, 19)), B(ep, (ep.numel(),))])
    assert lib.PupperResetFfi(C.addressof(fr)) is None, x.errors
    st = abi.PupperState()
    lib.pupper_state_blob_bind(C.byref(cfg), n, C.c_void_p(state.data_ptr()), C.byref(st))
    H36 = cfg.observation_history * 36

    def blob_field(name):
        off = (getattr(st, name) - state.data_ptr()) // 4
        rows = rt._fields[name].shape[0] if name != "obs" else None
        return state[off: off + rows * stride].view(rows, stride)[:, :n] if rows else state[off: off + n * H36].view(n, H36)

    def compare(tag):
        torch.cuda.synchronize()
        for name in ("qpos", "qvel", "qacc_warmstart", "last_act", "action_buffer", "imu_buffer", "command", "feet_air_time", "kick"):
            assert torch.equal(blob_field(name), rt.field(name)), (tag, name)
        assert torch.equal(blob_field("rng").view(torch.int32), rt.field("rng").view(torch.int32)), tag
        assert torch.equal(blob_field("obs"), rt.obs), tag
        assert torch.equal(reward, rt.reward) and torch.equal(done, rt.done) and torch.equal(metrics.view(n, 19), rt.metrics), tag

    compare("reset")
    state2, ep2 = torch.empty_like(state), torch.empty_like(ep)   # odd steps: results NOT aliased to the arguments
    for t in range(60):
        a = torch.from_numpy(common.actions(n, t)).to(dev)
        rt.step(a)
        aliased = t % 2 == 0
        s_out, e_out = (state, ep) if aliased else (state2, ep2)
        fr = x.frame([B(a, (n, 12)), B(state, (state.numel(),)), B(drb, (drb.numel(),)), B(ep, (ep.numel(),)), B(empty, (0,))],
                     [B(s_out, (state.numel(),)), B(reward, (n,)), B(done, (n,)), B(metrics, (n, 19)), B(e_out, (ep.numel(),))])
        assert lib.PupperStepFfi(C.addressof(fr)) is None, x.errors
        if not aliased:
            state.copy_(state2); ep.copy_(ep2)
        compare(f"step {t}")
    assert float(rt.done.sum()) >= 0 and float(rt.episode_field("totals")[0]) > 0   # episodes of 50 steps ended and auto-reset
    torch.cuda.synchronize()
    ta, tb = ep[-24:].cpu().numpy(), rt.episode_field("totals").cpu().numpy()  # atomicAdd sums: order-dependent rounding
    assert ta[0] == tb[0] and ta[2] == tb[2] and ta[22] == tb[22]
    np.testing.assert_allclose(ta, tb, rtol=1e-4, atol=1e-3)
    assert lib.pupper_ffi_unregister_model(0) == 0
    fr = x.frame([B(a, (n, 12)), B(state, (state.numel(),)), B(drb, (drb.numel(),)), B(ep, (ep.numel(),)), B(empty, (0,))],
                 [B(state, (state.numel(),)), B(reward, (n,)), B(done, (n,)), B(metrics, (n, 19)), B(ep, (ep.numel(),))])
    assert lib.PupperStepFfi(C.addressof(fr)) == 0xDEAD and "no model registered" in x.errors[-1][0]


@pytest.mark.gpu
def test_policy_and_rollout_through_the_ffi_handlers_match_the_runtime():
    """PupperPolicyFfi = pupper_policy_forward, PupperRolloutFfi = pupper_rollout on blobs: one custom call returns the
    [T, n, ...] trajectory and the advanced state / episode blobs, equal to what the ctypes / torch path produces."""
    torch = pytest.importorskip("torch")
    from pupperv3_mjx_b200 import rollout
    lib = _lib()
    env = common.make_env()
    env.set_episode_params(6, 1)
    cfg = env.env_cfg
    n, T = 136, 9
    dev = torch.device("cuda", 0)
    keys = common.env_keys(n)
    rt = runtime.EnvRuntime(env.model_desc, cfg, n, episode=True)
    d_keys = torch.from_numpy(keys.view(np.int32)).to(dev)
    rt.reset(d_keys)
    pol = rollout.PolicyMLP.random(env.observation_size, impl="cuda", precision=3, seed=7)
    w = env.observation_size
    assert lib.pupper_ffi_register_model(0, rt._model, C.byref(cfg)) == 0
    assert lib.pupper_ffi_register_policy(0, pol._kernel._handle, w, 12) == 0
    zeros = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
    x = FakeXla(stream=torch.cuda.current_stream().cuda_stream)
    B = lambda t, dims, dt=F32: x.buf(t.data_ptr() if t.numel() else None, dims, dt)
    # the policy alone
    obs = torch.randn((n, w), device=dev)
    act = zeros(n, 12)
    assert lib.PupperPolicyFfi(C.addressof(x.frame([B(obs, (n, w))], [B(act, (n, 12))]))) is None, x.errors
    torch.cuda.synchronize()
    assert torch.equal(act, pol(obs))
    assert lib.PupperPolicyFfi(C.addressof(x.frame([B(obs[:, :8].contiguous(), (n, 8))], [B(act, (n, 12))]))) == 0xDEAD  # wrong width
    # reset into blobs, then one unroll as ONE custom call against the runtime's rollout
    state = zeros(int(lib.pupper_state_blob_words(C.byref(cfg), n)))
    ep = zeros(int(lib.pupper_episode_blob_words(C.byref(cfg), n)))
    reward, done, metrics, empty = zeros(n), zeros(n), zeros(n * 19), zeros(0)
    fr = x.frame([B(d_keys, (n, 2), U32), B(empty, (0,)), B(empty, (0,))],
                 [B(state, (state.numel(),)), B(reward, (n,)), B(done, (n,)), B(metrics, (n, 19)), B(ep, (ep.numel(),))])
    assert lib.PupperResetFfi(C.addressof(fr)) is None, x.errors
    t_obs, t_act, t_rew, t_done = zeros(T, n, w), zeros(T, n, 12), zeros(T, n), zeros(T, n)
    fr = x.frame([B(state, (state.numel(),)), B(empty, (0,)), B(ep, (ep.numel(),))],
                 [B(state, (state.numel(),)), B(ep, (ep.numel(),)), B(t_obs, (T, n, w)), B(t_act, (T, n, 12)), B(t_rew, (T, n)), B(t_done, (T, n)),
                  B(metrics, (n, 19))])
    assert lib.PupperRolloutFfi(C.addressof(fr)) is None, x.errors
    r_obs, r_act, r_rew, r_done = zeros(T, n, w), zeros(T, n, 12), zeros(T, n), zeros(T, n)
    rt.rollout(pol._kernel, r_obs, r_act, r_rew, r_done)
    torch.cuda.synchronize()
    for a, b, name in ((t_obs, r_obs, "obs"), (t_act, r_act, "action"), (t_rew, r_rew, "reward"), (t_done, r_done, "done")):
        assert torch.equal(a, b), name
    assert float(t_done.sum()) > 0  # episodes of 6 steps ended inside the unroll
    st = abi.PupperState()
    lib.pupper_state_blob_bind(C.byref(cfg), n, C.c_void_p(state.data_ptr()), C.byref(st))
    off = (st.obs - state.data_ptr()) // 4
    assert torch.equal(state[off: off + n * w].view(n, w), rt.obs)
    assert torch.equal(metrics.view(n, 19), rt.metrics)
    assert lib.pupper_ffi_unregister_policy(0) == 0 and lib.pupper_ffi_unregister_model(0) == 0
