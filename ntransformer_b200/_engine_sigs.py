"""ctypes signatures of the native engine entry points (include/ntransformer.h, include/nt_b200_engine.h)."""
import ctypes as C

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class ModelConfigC(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("vocab_size", "hidden_size", "intermediate_size", "n_layers", "n_heads", "n_kv_heads",
                                       "head_dim", "max_seq_len")] + [("norm_eps", C.c_float), ("rope_theta", C.c_float),
                                                                      ("bos_token_id", C.c_int), ("eos_token_id", C.c_int)]


ENGINE_SIGNATURES = {
    # include/ntransformer.h
    "nt_engine_create": (_vp, []),
    "nt_engine_destroy": (None, [_vp]),
    "nt_engine_load": (_i, [_vp, C.c_char_p]),
    "nt_engine_generate": (_vp, [_vp, C.c_char_p, _i, _f, _i, _f]),
    "nt_free": (None, [_vp]),
    "nt_engine_vocab_size": (_i, [_vp]),
    "nt_engine_n_layers": (_i, [_vp]),
    "nt_engine_hidden_size": (_i, [_vp]),
    # include/nt_b200_engine.h
    "nt_model_load_gguf": (_vp, [C.c_char_p, _i, _i, _i]),
    "nt_model_create": (_vp, [C.POINTER(ModelConfigC), _i, _i]),
    "nt_model_set_tensor": (_i, [_vp, C.c_char_p, _vp, _i, _sz]),
    "nt_model_finalize": (_i, [_vp]),
    "nt_model_free": (None, [_vp]),
    "nt_model_get_config": (_i, [_vp, C.POINTER(ModelConfigC)]),
    "nt_model_forward": (_i, [_vp, _vp, _i, _i, _vp]),
    "nt_model_forward_async": (_i, [_vp, _vp, _i, _i]),
    "nt_model_sync": (_i, [_vp]),
    "nt_model_logits_device": (_vp, [_vp]),
    "nt_model_stream": (_vp, [_vp]),
    "nt_model_argmax": (_i, [_vp]),
    "nt_model_clear_kv": (None, [_vp]),
    "nt_model_use_graph": (None, [_vp, _i]),
    "nt_model_set_prefill_min_tokens": (None, [_vp, _i]),
    "nt_model_bytes_per_token": (C.c_ulonglong, [_vp, _i]),
    "nt_model_sample": (_i, [_vp, _f, _i, _f, _f, _vp, _i, _f]),
    "nt_sampler_uniform": (_f, [C.c_uint64, _i]),
    "nt_model_use_megakernel": (None, [_vp, _i]),
    "nt_model_tp_exchange": (_i, [_vp]),
    "nt_model_load_seconds": (C.c_double, [_vp]),
    "nt_model_megakernel_active": (_i, [_vp]),
    "nt_model_megakernel_plan": (_i, [_vp, _vp, _i]),
    "nt_model_megakernel_trace": (None, [_vp, _i]),
    "nt_model_megakernel_trace_read": (C.c_longlong, [_vp, _vp, _sz]),
    "nt_model_debug_read": (C.c_longlong, [_vp, C.c_char_p, _vp, _sz]),
    "nt_mega_plan_selftest": (_i, [C.POINTER(ModelConfigC), _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _sz]),
    "nt_gguf_describe": (_i, [C.c_char_p, _vp, _sz]),
    "nt_tokenize": (_i, [C.c_char_p, C.c_char_p, _i, _vp, _i]),
    "nt_detokenize": (_i, [C.c_char_p, _vp, _i, _vp, _sz]),
    "nt_sample_token": (_i, [_vp, _i, _f, _i, _f, _f, _i, _vp, _i, C.c_uint64]),
    "nt_tp_unique_id": (_i, [_vp]),
    "nt_tp_init": (_i, [_vp, _vp, _i, _i]),
}


def apply(lib):
    for name, (res, args) in ENGINE_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
