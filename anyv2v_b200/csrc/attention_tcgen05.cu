// placeholder until the tcgen05 attention kernel lands
#include "host_util.cuh"
extern "C" int av2v_attn_pnp_f16(const av2v_attn_args*, av2v_stream_t) {
  return av2v::fail(AV2V_ENOSUP, "attention kernel not built");
}
