// placeholder until the AMaZE kernels land (next commit)
#include "ctx.h"
int art_amaze_dev(art_hp_ctx* ctx, int, int, unsigned, const float*, size_t, float*, float*, float*, size_t, double, int)
{
    return ctx->fail(ART_HP_ERR_UNSUPPORTED, "AMaZE kernels not built yet");
}
