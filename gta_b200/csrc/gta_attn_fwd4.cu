// Single-launch forward (gta_attn_fwd4.cuh): instantiations with the head layout taken from the call at run time, and the
// dispatcher.  gta_attn_fwd4_ct.cu holds the instantiations specialised for the shipped layouts.
#include "gta_attn_fwd4.cuh"

namespace gta {

int launch_attn_fwd_v3_ct(const GtaAttnParams& p, const AttnArgs& a, const Fused4Args& f, cudaStream_t st, bool* handled);

template <typename TIn, typename TOut>
static int launch4_d(const AttnArgs& a, const Fused4Args& f, const GtaAttnParams& p, cudaStream_t st) {
    switch (p.D) {
        case 32: return launch4_one<TIn, TOut, 32, void>(a, f, p, st);
        case 64: return launch4_one<TIn, TOut, 64, void>(a, f, p, st);
        case 96: return launch4_one<TIn, TOut, 96, void>(a, f, p, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED, "persistent pipeline supports head dims 32/64/96");
}

// fused = true: K/V rotation inside the launch (the product path); false: the workspace already holds K'/V' (staged by
// rotate_kv_kernel: GTA_FLAG_SKIP_STAGE, or the two-launch path kept for measurement).
int launch_attn_fwd_v3(const GtaAttnParams& p, bool fused, cudaStream_t st) {
    const AttnArgs a = make_attn_args(p);
    Fused4Args f;
    f.rot = make_rot_args_kv(p);
    f.nunits = p.B * p.H * a.ntiles_k;
    f.flags = fused ? reinterpret_cast<int*>(static_cast<uint8_t*>(p.workspace) + kv_flags_offset(p.B, p.H, p.Tk, p.D)) : nullptr;
    const bool ib = p.in_dtype == GTA_DTYPE_BF16, ob = p.out_dtype == GTA_DTYPE_BF16;
    if (ib && !p.debug_clocks && !(p.flags & GTA_FLAG_RUNTIME_LAYOUT)) {      // shipped layouts: staging code specialised at compile time
        bool handled = false;
        const int rc = launch_attn_fwd_v3_ct(p, a, f, st, &handled);
        if (handled) return rc;
    }
    if (ib && ob) return launch4_d<__nv_bfloat16, __nv_bfloat16>(a, f, p, st);
    if (ib && !ob) return launch4_d<__nv_bfloat16, float>(a, f, p, st);
    if (!ib && ob) return launch4_d<float, __nv_bfloat16>(a, f, p, st);
    return launch4_d<float, float>(a, f, p, st);
}

}  // namespace gta
