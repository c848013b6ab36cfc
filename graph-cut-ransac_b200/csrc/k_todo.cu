// placeholders for entry points whose kernels land later in the round (kept so the ABI is complete)
#include "gcr_internal.cuh"

extern "C" gcr_status gcr_refit(gcr_ctx *ctx, int, int, const uint32_t *, const uint32_t *, const uint32_t *, int,
                                const double *, int, double *, uint8_t *)
{
	if (!ctx)
		return GCR_ERR_ARG;
	GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_refit: kernel not built yet");
}
extern "C" gcr_status gcr_grid_build(gcr_ctx *ctx, int, const double *, uint64_t, uint64_t *, uint32_t *, uint32_t *,
                                     uint32_t *, uint64_t *)
{
	if (!ctx)
		return GCR_ERR_ARG;
	GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_grid_build: kernel not built yet");
}
