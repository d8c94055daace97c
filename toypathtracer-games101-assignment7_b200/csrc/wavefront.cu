// wavefront.cu — placeholder until the queue pipeline lands (next commit).
#include "tpt_internal.h"

int wavefront_render(TptScene*, const RenderArgs&, float*, float*, cudaStream_t, unsigned long long*) {
    tpt_set_error("wavefront pipeline not built yet");
    return TPT_ERR_INVALID;
}
void wavefront_destroy(TptScene*) {}
