// Placeholder translation unit for the tcgen05 flash-attention kernel (prefill / NAR); see DESIGN.md.
#include "common.cuh"

extern "C" int vb_attention_prefill_tc(const void* qkv, void* o, int B, int S, int H, int mask_mode, const int32_t* x_lens,
                                       const int32_t* kv_lens, void* stream) {
    (void)qkv; (void)o; (void)B; (void)S; (void)H; (void)mask_mode; (void)x_lens; (void)kv_lens; (void)stream;
    VB_REQUIRE(false, VB_ERR_UNSUPPORTED, "vb_attention_prefill_tc: not built in this revision");
}
