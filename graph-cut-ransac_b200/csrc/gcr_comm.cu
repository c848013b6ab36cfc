// gcr_comm.cu -- the hypothesis-sharded mode of SURVEY.md 8(e): ONE problem, its K samples split over the GPUs.
//
// The reference has no counterpart (it is single-threaded); what must be preserved across ranks is the running-best
// rule of GCRANSAC.h:440 with Score::operator< (scoring_function.h:63-66): strictly larger FP64 score wins, among equal
// scores the FIRST model (lowest global index).  Every rank scores its contiguous range of the counter-based sample
// stream in chunks, keeps its running winner as one 128-byte gcr_winner record on the device (k_select_exact's
// accumulate mode), the records are all-gathered over NVLink (ncclAllGather: 128 B per pair per rank) and every rank
// reduces them with the same symmetric rule, so all ranks hold the identical winner -- score, index, inlier count AND
// model -- after one collective and no host round trip.  (A 64-bit ncclMax key cannot carry an FP64 score plus a 64-bit
// index; the round-1 packed FP32 key decided near-ties in FP32.)
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a torch process this resolves to the copy torch already
// loaded, in a plain C / C++ host to the system library.  Only types come from <nccl.h>.
#include "gcr_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

namespace
{
struct NcclApi
{
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	char err[256] = {0};
};

NcclApi *nccl_api()
{
	static NcclApi api;
	static bool tried = false;
	if (tried)
		return api.handle ? &api : nullptr;
	tried = true;
	const char *names[] = {"libnccl.so.2", "libnccl.so"};
	for (const char *nm : names)
	{
		api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
		if (api.handle)
			break;
	}
	if (!api.handle)
	{
		snprintf(api.err, sizeof(api.err), "dlopen(libnccl.so.2) failed: %s", dlerror());
		return nullptr;
	}
	api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
	api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
	api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
	api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
	api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
	if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString)
	{
		snprintf(api.err, sizeof(api.err), "libnccl.so.2 lacks a required symbol");
		api.handle = nullptr;
		return nullptr;
	}
	return &api;
}
} // namespace

#define GCR_CHECK_NCCL(ctx, api, expr)                                                                         \
	do                                                                                                         \
	{                                                                                                          \
		ncclResult_t _r = (expr);                                                                              \
		if (_r != ncclSuccess)                                                                                 \
		{                                                                                                      \
			snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,            \
			         (api)->GetErrorString(_r));                                                               \
			return GCR_ERR_CUDA;                                                                               \
		}                                                                                                      \
	} while (0)

static_assert(sizeof(gcr_winner) == 128, "gcr_winner is exchanged as 128 raw bytes");
static_assert(sizeof(ncclUniqueId) == 128, "gcr_comm_unique_id hands out 128 bytes");

extern "C" gcr_status gcr_comm_unique_id(void *h_out_128)
{
	if (!h_out_128)
		return GCR_ERR_ARG;
	NcclApi *api = nccl_api();
	if (!api)
		return GCR_ERR_STATE;
	ncclUniqueId id;
	if (api->GetUniqueId(&id) != ncclSuccess)
		return GCR_ERR_CUDA;
	memcpy(h_out_128, &id, sizeof(id));
	return GCR_OK;
}

extern "C" gcr_status gcr_comm_init(gcr_ctx *ctx, const void *h_unique_id_128, int rank, int world)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!h_unique_id_128 || world < 1 || rank < 0 || rank >= world)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_comm_init: need a unique id and 0 <= rank < world (got %d / %d)", rank, world);
	if (ctx->nccl_comm)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_comm_init: the context already has a communicator");
	NcclApi *api = nccl_api();
	if (!api)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_comm_init: NCCL is not available (libnccl.so.2 could not be loaded)");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	ncclUniqueId id;
	memcpy(&id, h_unique_id_128, sizeof(id));
	ncclComm_t comm = nullptr;
	GCR_CHECK_NCCL(ctx, api, api->CommInitRank(&comm, world, id, rank));
	ctx->nccl_comm = comm;
	ctx->comm_rank = rank;
	ctx->comm_world = world;
	return GCR_OK;
}

extern "C" gcr_status gcr_comm_destroy(gcr_ctx *ctx)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (ctx->nccl_comm)
	{
		NcclApi *api = nccl_api();
		cudaSetDevice(ctx->device);
		cudaStreamSynchronize(ctx->stream);
		if (api)
			api->CommDestroy((ncclComm_t)ctx->nccl_comm);
		ctx->nccl_comm = nullptr;
	}
	ctx->comm_rank = 0;
	ctx->comm_world = 1;
	return GCR_OK;
}

// all [world][n_pairs] -> out [n_pairs]: larger FP64 score, then lower index; symmetric, so every rank computes the same
__global__ void k_winners_reduce(const gcr_winner *__restrict__ all, int world, int n_pairs, gcr_winner *__restrict__ out)
{
	const int pair = blockIdx.x * blockDim.x + threadIdx.x;
	if (pair >= n_pairs)
		return;
	int best = -1;
	double bs = 0.0;
	unsigned long long bi = 0ull;
	for (int r = 0; r < world; ++r)
	{
		const gcr_winner *w = &all[(size_t)r * n_pairs + pair];
		if (!w->valid)
			continue;
		if (best < 0 || w->score > bs || (w->score == bs && w->index < bi))
		{
			best = r;
			bs = w->score;
			bi = w->index;
		}
	}
	gcr_winner res;
	if (best >= 0)
		res = all[(size_t)best * n_pairs + pair];
	else
		memset(&res, 0, sizeof(res));
	out[pair] = res;
}

extern "C" gcr_status gcr_winners_reduce(gcr_ctx *ctx, const gcr_winner *d_all, int world, gcr_winner *d_out)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_all || !d_out || world < 1 || ctx->n_pairs <= 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_winners_reduce: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	k_winners_reduce<<<(ctx->n_pairs + 127) / 128, 128, 0, ctx->stream>>>(d_all, world, ctx->n_pairs, d_out);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

// contiguous, balanced [start, stop) of `total` units for `rank` (the first total % world ranks get one more)
static void shard_range(uint64_t total, int rank, int world, uint64_t *lo, uint64_t *hi)
{
	const uint64_t base = total / world, extra = total % world;
	*lo = (uint64_t)rank * base + ((uint64_t)rank < extra ? (uint64_t)rank : extra);
	*hi = *lo + base + ((uint64_t)rank < extra ? 1 : 0);
}

extern "C" gcr_status gcr_round_sharded(gcr_ctx *ctx, int problem, int scoring, int sampler, uint64_t first_index,
                                        uint64_t total_samples, uint64_t seed, double threshold, int flags, int chunk_samples,
                                        gcr_winner *d_winners, uint64_t *evals_out)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_winners || chunk_samples <= 0 || ctx->n_pairs <= 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round_sharded: need d_winners, chunk_samples > 0 and uploaded points");
	const int mm = gcr_max_models(problem);
	if (mm == 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round_sharded: unknown problem %d", problem);
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const int world = ctx->nccl_comm ? ctx->comm_world : 1, rank = ctx->nccl_comm ? ctx->comm_rank : 0;
	const size_t P = (size_t)ctx->n_pairs;
	if ((size_t)world * P > ctx->gather_capacity)
	{
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_gather);
		ctx->d_gather = nullptr;
		ctx->gather_capacity = 0;
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_gather, (size_t)world * P * sizeof(gcr_winner)));
		ctx->gather_capacity = (size_t)world * P;
	}
	gcr_winner *mine = ctx->d_gather + (size_t)rank * P; // in place: NCCL's all-gather takes it from here
	uint64_t lo, hi, evals = 0;
	shard_range(total_samples, rank, world, &lo, &hi);
	if (lo == hi)
		GCR_CHECK_CUDA(ctx, cudaMemsetAsync(mine, 0, P * sizeof(gcr_winner), ctx->stream)); // valid = 0
	for (uint64_t c0 = lo; c0 < hi; c0 += (uint64_t)chunk_samples)
	{
		const int count = (int)(hi - c0 < (uint64_t)chunk_samples ? hi - c0 : (uint64_t)chunk_samples);
		uint64_t ev = 0;
		gcr_status st = gcr_round_winners(ctx, problem, scoring, sampler, first_index + c0, count, seed, threshold, flags,
		                                  c0 * (uint64_t)mm, c0 != lo, mine, &ev);
		if (st)
			return st;
		evals += ev;
	}
	if (world > 1)
	{
		NcclApi *api = nccl_api();
		GCR_CHECK_NCCL(ctx, api, api->AllGather(mine, ctx->d_gather, P * sizeof(gcr_winner), ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
		gcr_status st = gcr_winners_reduce(ctx, ctx->d_gather, world, d_winners);
		if (st)
			return st;
	}
	else
		GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(d_winners, mine, P * sizeof(gcr_winner), cudaMemcpyDeviceToDevice, ctx->stream));
	if (evals_out)
		*evals_out = evals;
	return GCR_OK;
}
