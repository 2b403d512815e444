// placeholder, replaced below
#include "gf_common.cuh"
namespace gf {
bool tc_supported(const Layout&, const gf_attn_desc*) { return false; }
int token_pass_tc(const Layout&, const gf_attn_desc*, const float*, float*, float*, float*, cudaStream_t) { set_error("tc path not built"); return GF_ERR_UNSUPPORTED; }
}
