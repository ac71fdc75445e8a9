#include "host.h"
namespace lft { int configure_up(){return 0;} }
using namespace lft;
extern "C" {
#define UNIMPL return fail(LFT_ERR_STATE, "not implemented yet")
int lft_forward(lft_handle*, const float*, float*, int32_t, int32_t, void*, size_t, void*) { UNIMPL; }
int lft_lf_num_patches(int32_t, int32_t, int32_t*, int32_t*) { UNIMPL; }
int lft_forward_lf(lft_handle*, const float*, int32_t, int32_t, int32_t, int32_t, float*, void*, size_t, void*) { UNIMPL; }
int lft_integrate(lft_handle*, const float*, int32_t, int32_t, int32_t, int32_t, float*, void*) { UNIMPL; }
int lft_divide(lft_handle*, const float*, int32_t, int32_t, int32_t, int32_t, float*, void*) { UNIMPL; }
int lft_stage_upsample(lft_handle*, const float*, const float*, float*, int32_t, int32_t, void*, size_t, void*) { UNIMPL; }
}
