// barneshut.cu — placeholder while the tree kernels land (replaced in the next commit).
#include "common.cuh"
namespace pcuda {
void tree_free(pcuda_ctx *, pcuda_tree *) {}
}  // namespace pcuda
using namespace pcuda;
#define NI(ctx) return fail(ctx, PCUDA_ERR_NOT_INITIALISED, "Barnes-Hut not built yet")
extern "C" {
int pcuda_barneshut_f32x3(pcuda_ctx *c, const float *, size_t, const float *, size_t, float, float, int, float *) { NI(c); }
int pcuda_barneshut_f32x2(pcuda_ctx *c, const float *, size_t, const float *, size_t, float, float, int, float *) { NI(c); }
int pcuda_barneshut_f32x3_dev(pcuda_ctx *c, const float *, size_t, const float *, size_t, float, float, int, float *) { NI(c); }
int pcuda_barneshut_f32x2_dev(pcuda_ctx *c, const float *, size_t, const float *, size_t, float, float, int, float *) { NI(c); }
int pcuda_tree_build_f32(pcuda_ctx *c, uint32_t, const float *, size_t, pcuda_tree **) { NI(c); }
int pcuda_tree_info_get(const pcuda_tree *, pcuda_tree_info *) { return PCUDA_ERR_NOT_INITIALISED; }
int pcuda_tree_read(pcuda_ctx *c, const pcuda_tree *, int, void *, size_t) { NI(c); }
int pcuda_tree_traverse_f32(pcuda_ctx *c, const pcuda_tree *, const float *, size_t, float, float, int, float *) { NI(c); }
int pcuda_tree_last_counters(pcuda_ctx *c, uint64_t *) { NI(c); }
void pcuda_tree_destroy(pcuda_ctx *, pcuda_tree *) {}
}
