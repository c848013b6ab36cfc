// gcr_batch_engine_impl.h -- the lock-step multi-pair driver (see gcr_batch_engine.h); included by gcr_driver.cpp after the
// Driver class, whose control flow (the sequential replay of GCRANSAC::run, GCRANSAC.h:202-627) it reuses unchanged.
//
//   Engine          one gcr_ctx with S resident image pairs ("seats"), one dispatcher thread.  S worker threads each run
//                   the Driver of one pair at a time over a BatchedBackend whose operations are REQUESTS: the worker
//                   posts one and sleeps.  The dispatcher takes whatever is pending, groups it by kind (+ the launch
//                   parameters that must agree) and executes every group as ONE launch sequence over all the seats that
//                   asked, through the per-item-pair entry points of include/gcr/gcr.h, with one synchronise per group.
//                   Requests that arrive while the GPU works form the next super-step: the batches grow by themselves
//                   when the dispatcher is the bottleneck.
//   results         land in pinned staging areas; every requester copies its own rows out after it wakes (the copies run
//                   on the host cores in parallel, the dispatcher only waits for a staging area's readers before it
//                   reuses it).
//   host-only work  between two requests -- the reference's Boykov-Kolmogorov min-cut above all -- runs in the worker
//                   threads, i.e. fanned out over the host cores.
//   several engines (contexts, streams, dispatchers) share one queue of pairs: one engine's host-side staging overlaps
//                   the other's GPU work.
#ifndef GCR_BATCH_ENGINE_IMPL_H_INCLUDED
#define GCR_BATCH_ENGINE_IMPL_H_INCLUDED

#include "gcr_batch_engine.h"

#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>

#include <deque>

#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <ucontext.h>
#include <unistd.h>

namespace
{
	using gcr_host::DeviceBuffer;

	const uint64_t kPairSeedStride = 0x9E3779B97F4A7C15ull; // GCR_PAIR_SEED_STRIDE of the sampling kernels

	struct PinnedBuffer
	{
		gcr_ctx *ctx = nullptr;
		void *p = nullptr;
		size_t bytes = 0;
		~PinnedBuffer() { if (p) gcr_host_free_pinned(ctx, p); }
		bool ensure(gcr_ctx *c, size_t need)
		{
			ctx = c;
			if (need <= bytes)
				return true;
			if (p) gcr_host_free_pinned(ctx, p);
			p = nullptr;
			bytes = 0;
			// grow in powers of two from 1 MiB: the groups of consecutive super-steps differ in size, and allocating pinned memory
			// stalls every stream of the process
			size_t want = (size_t)1 << 20;
			while (want < need)
				want <<= 1;
			if (gcr_host_alloc_pinned(ctx, want, &p) != GCR_OK)
				return false;
			bytes = want;
			return true;
		}
		template <class T> T *at(size_t off) const { return reinterpret_cast<T *>(static_cast<char *>(p) + off); }
	};

	// a pinned result area that worker threads read after the dispatcher has filled it
	struct Staging : PinnedBuffer
	{
		std::atomic<int> readers{0};
		void wait_free()
		{
			while (readers.load(std::memory_order_acquire) > 0)
				std::this_thread::yield();
		}
	};

	// inputs of one group: arrays carved out of ONE pinned block with a device twin of the same layout -> one H2D copy
	struct InBlock
	{
		PinnedBuffer h;
		DeviceBuffer d;
		size_t used = 0;
		void reset() { used = 0; }
		size_t carve(size_t bytes)
		{
			const size_t o = used;
			used += (bytes + 255) & ~(size_t)255;
			return o;
		}
		bool commit(gcr_ctx *ctx) { return h.ensure(ctx, used ? used : 256) && d.ensure(ctx, used ? used : 256); }
		template <class T> T *host(size_t off) const { return h.at<T>(off); }
		template <class T> T *dev(size_t off) const { return reinterpret_cast<T *>(static_cast<char *>(d.p) + off); }
		bool upload(gcr_ctx *ctx, size_t from, size_t to) const
		{
			return to <= from || gcr_memcpy_pinned_async(ctx, static_cast<char *>(d.p) + from, static_cast<const char *>(h.p) + from, to - from, 1) == GCR_OK;
		}
	};

	// ---- fibers: the Driver of one pair is sequential, blocking code (it posts a request and needs the answer); a pair in
	// flight is a user-level context on a stack of its own, and a few OS threads (about one per core) run whichever pairs
	// are runnable.  Posting a request switches back to the thread's scheduler; the dispatcher's completion makes the
	// fiber runnable again, on any thread.  No kernel wake-up per request, no more runnable OS threads than cores.
	struct Fiber
	{
		ucontext_t uc;
		void *stack = nullptr;
		size_t stack_bytes = 0;
		std::function<void()> body;
		bool finished = false;
		// RUNNING: on a thread (or still switching out); PARKED: switched out, waiting; READY: the completion arrived before the
		// switch-out finished.  Whoever sees the other side's mark enqueues the fiber: exactly one of the two does.
		enum { RUNNING = 0, PARKED = 1, READY = 2 };
		std::atomic<int> state{RUNNING};
		~Fiber() { if (stack) munmap(stack, stack_bytes); }
	};
	struct RunQueue
	{
		std::mutex mu;
		std::condition_variable cv;
		std::deque<Fiber *> q;
		int live = 0; // fibers not finished yet
		void push(Fiber *f)
		{
			std::lock_guard<std::mutex> lock(mu);
			q.push_back(f);
			cv.notify_one();
		}
		Fiber *pop_wait() // null: every fiber has finished
		{
			std::unique_lock<std::mutex> lock(mu);
			cv.wait(lock, [&] { return !q.empty() || live == 0; });
			if (q.empty())
				return nullptr;
			Fiber *f = q.front();
			q.pop_front();
			return f;
		}
		void finished_one()
		{
			std::lock_guard<std::mutex> lock(mu);
			if (--live == 0)
				cv.notify_all();
		}
	};
	thread_local ucontext_t *tl_scheduler = nullptr; // the OS thread's own context (its fiber loop)
	thread_local Fiber *tl_fiber = nullptr;          // the fiber this thread is running
	// (never inlined: after a switch the fiber may be on another thread, the thread-local addresses must be looked up again)
	__attribute__((noinline)) ucontext_t *current_scheduler() { return tl_scheduler; }
	__attribute__((noinline)) Fiber *current_fiber() { return tl_fiber; }
	void fiber_trampoline(unsigned lo, unsigned hi)
	{
		Fiber *f = reinterpret_cast<Fiber *>(((uintptr_t)hi << 32) | (uintptr_t)lo);
		f->body();
		f->finished = true;
		swapcontext(&f->uc, current_scheduler());
	}
	bool fiber_start(Fiber *f, size_t stack_bytes)
	{
		f->stack_bytes = stack_bytes;
		f->stack = mmap(nullptr, stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK | MAP_NORESERVE, -1, 0);
		if (f->stack == MAP_FAILED)
		{
			f->stack = nullptr;
			return false;
		}
		getcontext(&f->uc);
		f->uc.uc_stack.ss_sp = f->stack;
		f->uc.uc_stack.ss_size = stack_bytes;
		f->uc.uc_link = nullptr;
		const uintptr_t p = reinterpret_cast<uintptr_t>(f);
		makecontext(&f->uc, (void (*)())fiber_trampoline, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
		return true;
	}
	// the loop of one OS worker thread
	void fiber_thread(RunQueue *rq)
	{
		// the dispatcher threads must get a core the moment requests are pending; the fiber threads (host min-cuts) yield to
		// them ($GCR_BATCH_WORKER_NICE, default 5; raising one's own niceness needs no privilege)
		static const int worker_nice = getenv("GCR_BATCH_WORKER_NICE") ? atoi(getenv("GCR_BATCH_WORKER_NICE")) : 5;
		if (worker_nice > 0)
			setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), worker_nice);
		ucontext_t self;
		tl_scheduler = &self;
		while (Fiber *f = rq->pop_wait())
		{
			tl_fiber = f;
			swapcontext(&self, &f->uc); // runs the fiber until it posts a request or finishes
			tl_fiber = nullptr;
			if (f->finished)
			{
				rq->finished_one();
				continue;
			}
			if (f->state.exchange(Fiber::PARKED, std::memory_order_acq_rel) == Fiber::READY) // completed while it switched out
			{
				f->state.store(Fiber::RUNNING, std::memory_order_release);
				rq->push(f);
			}
		}
	}

	struct Request
	{
		enum Kind { UPLOAD = 0, BATCH, EXACT, LABEL, REFIT, LO, FETCH, SERIAL, KINDS };
		Kind kind = UPLOAD;
		int seat = 0;
		// UPLOAD
		const double *pts = nullptr;
		// BATCH / FETCH
		const BatchSpec *bspec = nullptr;
		BatchOut *bout = nullptr;
		const std::vector<size_t> *slots = nullptr;
		// EXACT (models: [C][ml]) / LABEL (models: one model)
		const double *models = nullptr;
		size_t C = 0;
		int scoring = GCR_MSAC;
		double threshold = 0.0, lambda = 0.0;
		double *o_score = nullptr;
		uint32_t *o_count = nullptr;
		uint32_t *o_idx = nullptr; // row c at o_idx + c * n (the first count[c] entries are written)
		double *o_res = nullptr;   // [n], single-model requests only
		std::vector<double> *o_unary = nullptr;
		// REFIT
		int rp = 0, ba_frame = 0, weight_mode = GCR_WEIGHT_NONE;
		const uint32_t *flat = nullptr;
		const std::vector<uint32_t> *offsets = nullptr, *sizes = nullptr;
		const double *weights = nullptr;
		size_t n_weights = 0;
		std::vector<double> *o_models = nullptr;
		std::vector<uint8_t> *o_ok = nullptr;
		// LO
		const LoSpec *lo = nullptr;
		std::vector<float> *o_scores32 = nullptr;
		// SERIAL: an operation the engine does not batch (once per pair at most), run for this seat on the dispatcher thread
		std::function<bool(DirectBackend &)> fn;
		// completion
		Fiber *fiber = nullptr;
		bool done = false, ok = true;
		std::string error;
		Staging *staging = nullptr;     // the requester copies its results out of it (deliver) and then releases it
		std::function<void()> deliver;
	};

	class Engine
	{
	public:
		gcr_ctx *ctx = nullptr;
		int S = 0, n = 0, cols = 4, problem = 0, ml = 9, device = 0;
		std::string error;

		std::mutex mu;
		std::condition_variable cv_dispatch;
		RunQueue *run_queue = nullptr; // of the call in progress
		std::vector<Request *> pending;
		int live_workers = 0;
		std::thread dispatcher;
		// statistics
		uint64_t super_steps = 0, groups = 0, requests = 0;
		double kind_seconds[Request::KINDS] = {0}, idle_seconds = 0, sync_seconds = 0, complete_seconds = 0;
		uint64_t kind_requests[Request::KINDS] = {0}, kind_groups[Request::KINDS] = {0};

		~Engine()
		{
			// every device / pinned block goes back before the context it came from
			classes.clear();
			for (auto &pool : slots_of_kind)
				pool.clear();
			issued.clear();
			for (DeviceBuffer *b : {&lo_models, &lo_ok, &lo_scores})
				if (b->p)
				{
					gcr_free(ctx, b->p);
					b->p = nullptr;
					b->bytes = 0;
				}
			if (ctx) gcr_ctx_destroy(ctx);
		}

		bool create(int device_, int problem_, int n_, int seats, const double *first_points)
		{
			device = device_;
			problem = problem_;
			n = n_;
			S = seats;
			ml = gcr_model_len(problem);
			cols = problem == GCR_PNP ? 5 : 4;
			if (gcr_ctx_create(device, &ctx) != GCR_OK)
			{
				error = "gcr_ctx_create failed";
				return false;
			}
			return gcr_points_upload(ctx, first_points, n, cols, S) == GCR_OK || fail("gcr_points_upload");
		}

		// ---- worker side ----
		bool submit(Request &r)
		{
			Fiber *f = current_fiber();
			r.fiber = f;
			{
				std::lock_guard<std::mutex> lock(mu);
				pending.push_back(&r);
				cv_dispatch.notify_one();
			}
			swapcontext(&f->uc, current_scheduler()); // back when the dispatcher has completed the request (on any thread)
			if (r.ok && r.deliver)
				r.deliver();
			if (r.staging)
				r.staging->readers.fetch_sub(1, std::memory_order_release);
			return r.ok;
		}
		void worker_started()
		{
			std::lock_guard<std::mutex> lock(mu);
			++live_workers;
		}
		void worker_done()
		{
			std::lock_guard<std::mutex> lock(mu);
			--live_workers;
			cv_dispatch.notify_one();
		}

		// ---- dispatcher ----
		void run_dispatcher()
		{
			std::vector<Request *> batch;
			for (;;)
			{
				{
					const auto t0 = std::chrono::steady_clock::now();
					std::unique_lock<std::mutex> lock(mu);
					cv_dispatch.wait(lock, [&] { return !pending.empty() || live_workers == 0; });
					idle_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
					if (pending.empty() && live_workers == 0)
						return;
					batch.swap(pending);
				}
				++super_steps;
				requests += batch.size();
				std::vector<Request *> by_kind[Request::KINDS];
				for (Request *r : batch)
					by_kind[r->kind].push_back(r);
				batch.clear();
				// every group of every kind is queued on the stream, ONE synchronise delivers them all
				issued.clear();
				for (int k = 0; k < Request::SERIAL; ++k)
					if (!by_kind[k].empty())
					{
						const auto t0 = std::chrono::steady_clock::now();
						const uint64_t g0 = groups;
						kind_requests[k] += by_kind[k].size();
						execute_kind((Request::Kind)k, by_kind[k]);
						kind_groups[k] += groups - g0;
						kind_seconds[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
					}
				if (!issued.empty())
				{
					const auto t0 = std::chrono::steady_clock::now();
					const bool synced = gcr_sync(ctx) == GCR_OK;
					if (!synced)
						fail("gcr_sync");
					sync_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
					const auto t1 = std::chrono::steady_clock::now();
					for (Issued &rec : issued)
					{
						bool ok = rec.ok && synced;
						bool completed = false;
						if (ok && rec.post)
							ok = rec.post(completed);
						if (!completed)
							complete(rec.group, ok);
						rec.slot->busy = false; // (its staging area stays reserved until the readers are done)
					}
					complete_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
				}
				if (!by_kind[Request::SERIAL].empty())
				{
					const auto t0 = std::chrono::steady_clock::now();
					kind_requests[Request::SERIAL] += by_kind[Request::SERIAL].size();
					kind_groups[Request::SERIAL] += by_kind[Request::SERIAL].size();
					execute_kind(Request::SERIAL, by_kind[Request::SERIAL]);
					kind_seconds[Request::SERIAL] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
				}
			}
		}

	private:
		struct BatchClass // the device side of one batch size: [S] rows of samples / model slots / flags / scores / counts
		{
			DeviceBuffer samples, models, valid, scores, counts;
		};
		std::map<int, std::unique_ptr<BatchClass>> classes;
		std::vector<int> seat_class; // batch size the seat's current batch lives in
		// what one group needs between its launch and its delivery: pinned inputs + device twin, device outputs, a pinned
		// result area.  A slot is reused once its group is complete and the last requester has copied its rows out.
		struct Slot
		{
			InBlock in;
			DeviceBuffer d_a, d_b, d_c, d_d;
			Staging st;
			bool busy = false;
		};
		// one pool per kind of request: a slot's buffers then see groups of similar size and stop regrowing after a call or
		// two (pinned reallocations stall every stream: with one shared pool the passes of a call varied 2-8 x)
		std::vector<std::unique_ptr<Slot>> slots_of_kind[Request::KINDS];
		Request::Kind acquiring = Request::UPLOAD;
		struct Issued
		{
			std::vector<Request *> group;
			Slot *slot = nullptr;
			bool ok = true;
			std::function<bool(bool &completed)> post; // after the synchronise (BATCH: the overflow check)
		};
		std::vector<Issued> issued;
		Slot &acquire()
		{
			std::vector<std::unique_ptr<Slot>> &slots = slots_of_kind[acquiring];
			// A slot is free again once the last requester of its group has copied its rows out, i.e. once those fibers have
			// run.  A pool that simply grew whenever nothing was free kept allocating pinned areas (each a device-wide stall)
			// for as long as the fiber threads were busy with min-cuts: at most `kMaxSlots` per kind, then wait for one.
			const size_t kMaxSlots = 12;
			for (;;)
			{
				for (auto &sl : slots)
					if (!sl->busy && sl->st.readers.load(std::memory_order_acquire) == 0)
					{
						sl->busy = true;
						return *sl;
					}
				bool waiting_for_readers = false; // (slots of THIS super-step's groups stay busy until it completes: never wait for those)
				for (auto &sl : slots)
					waiting_for_readers |= !sl->busy;
				if (slots.size() < kMaxSlots || !waiting_for_readers)
					break;
				std::this_thread::yield();
			}
			slots.emplace_back(new Slot());
			slots.back()->busy = true;
			return *slots.back();
		}

		bool fail(const char *what)
		{
			error = std::string(what) + ": " + gcr_last_error(ctx);
			return false;
		}
		void complete(const std::vector<Request *> &group, bool ok)
		{
			for (Request *r : group)
			{
				r->ok = ok;
				if (!ok)
				{
					r->error = error;
					if (r->staging)
					{
						r->staging->readers.fetch_sub(1, std::memory_order_release);
						r->staging = nullptr;
					}
				}
				r->done = true;
				Fiber *f = r->fiber;
				if (f->state.exchange(Fiber::READY, std::memory_order_acq_rel) == Fiber::PARKED) // (else its thread enqueues it)
				{
					f->state.store(Fiber::RUNNING, std::memory_order_release);
					run_queue->push(f);
				}
			}
		}
		void attach(Request *r, Staging &st)
		{
			r->staging = &st;
			st.readers.fetch_add(1, std::memory_order_acq_rel);
		}

		// split `all` into groups whose members agree on `key`; each group is queued on the stream with a slot of its own
		template <class Key> void grouped(const std::vector<Request *> &all, Key key, const std::function<bool(std::vector<Request *> &, Issued &)> &run)
		{
			std::vector<Request *> rest(all), group, other;
			while (!rest.empty())
			{
				group.clear();
				other.clear();
				const auto k0 = key(rest[0]);
				for (Request *r : rest)
					(key(r) == k0 ? group : other).push_back(r);
				++groups;
				issued.emplace_back();
				Issued &rec = issued.back();
				rec.group = group;
				rec.slot = &acquire();
				rec.ok = run(group, rec);
				rest.swap(other);
			}
		}

		void execute_kind(Request::Kind kind, std::vector<Request *> &reqs)
		{
			acquiring = kind;
			switch (kind)
			{
			case Request::UPLOAD:
				grouped(reqs, [](Request *) { return 0; }, [&](std::vector<Request *> &g, Issued &rec) { return do_upload(g, *rec.slot); });
				break;
			case Request::BATCH:
				grouped(reqs,
				        [](Request *r) {
					        const BatchSpec &s = *r->bspec;
					        return std::make_tuple(s.B, s.m, s.mm, s.explicit_samples ? 2 : s.sampler, s.scoring, s.threshold, s.want_samples, s.variant);
				        },
				        [&](std::vector<Request *> &g, Issued &rec) { return do_batch(g, g[0]->bspec->B, rec); });
				break;
			case Request::EXACT:
				grouped(reqs, [](Request *r) { return std::make_tuple(r->scoring, r->threshold); },
				        [&](std::vector<Request *> &g, Issued &rec) { return do_exact(g, *rec.slot); });
				break;
			case Request::LABEL:
				grouped(reqs, [](Request *r) { return std::make_tuple(r->threshold, r->lambda); },
				        [&](std::vector<Request *> &g, Issued &rec) { return do_label(g, *rec.slot); });
				break;
			case Request::REFIT:
				grouped(reqs,
				        [](Request *r) {
					        return std::make_tuple(r->rp, r->ba_frame, r->weights ? r->weight_mode : (int)GCR_WEIGHT_NONE, r->weights != nullptr,
					                               r->sizes->size() == 1);
				        },
				        [&](std::vector<Request *> &g, Issued &rec) { return do_refit(g, *rec.slot); });
				break;
			case Request::LO:
				grouped(reqs,
				        [](Request *r) {
					        const LoSpec &s = *r->lo;
					        return std::make_tuple(s.sample_pool, s.refit_problem, s.ba_frame, s.rm, s.scoring, s.threshold, s.sample_pool ? s.trials : 1);
				        },
				        [&](std::vector<Request *> &g, Issued &rec) { return do_lo(g, *rec.slot); });
				break;
			case Request::FETCH:
				grouped(reqs, [&](Request *r) { return seat_class.empty() ? 0 : seat_class[r->seat]; },
				        [&](std::vector<Request *> &g, Issued &rec) { return do_fetch(g, *rec.slot); });
				break;
			case Request::SERIAL:
				for (Request *r : reqs)
				{
					DirectBackend direct(ctx, r->seat);
					direct.problem = problem;
					direct.n = n;
					direct.ml = ml;
					direct.cols = cols;
					const bool ok = r->fn(direct);
					if (!ok)
						error = direct.error;
					++groups;
					complete(std::vector<Request *>(1, r), ok);
				}
				break;
			default:
				break;
			}
		}

		// ---- UPLOAD: the pair of every requesting seat replaced in place (frames / FP32 copies re-derived on the device) ----
		bool do_upload(std::vector<Request *> &g, Slot &sl)
		{
			InBlock &in = sl.in;
			const size_t count = g.size(), row = (size_t)n * cols;
			in.reset();
			const size_t o_seats = in.carve(count * sizeof(int32_t)), o_pts = in.carve(count * row * sizeof(double));
			if (!in.commit(ctx))
				return fail("pinned / device allocation");
			for (size_t i = 0; i < count; ++i)
			{
				in.host<int32_t>(o_seats)[i] = g[i]->seat;
				memcpy(in.host<double>(o_pts) + i * row, g[i]->pts, row * sizeof(double));
			}
			// (queued: the pinned rows belong to this group's slot until the super-step's synchronise)
			return gcr_points_replace_pairs(ctx, in.host<int32_t>(o_seats), (int)count, in.host<double>(o_pts)) == GCR_OK || fail("gcr_points_replace_pairs");
		}

		BatchClass *batch_class(int B, int m, int mm)
		{
			std::unique_ptr<BatchClass> &c = classes[B];
			if (!c)
				c.reset(new BatchClass());
			const size_t K = (size_t)B * mm, SK = (size_t)S * K;
			if (!c->samples.ensure(ctx, (size_t)S * B * m * sizeof(uint32_t)) || !c->models.ensure(ctx, SK * ml * sizeof(double)) ||
			    !c->valid.ensure(ctx, SK) || !c->scores.ensure(ctx, SK * sizeof(float)) || !c->counts.ensure(ctx, SK * sizeof(uint32_t)))
				return nullptr;
			return c.get();
		}

		// ---- BATCH: samples [first, first + B) of every requesting seat -> models, flags, FP32 scores, exact counts ----
		// An overflowing FP64 re-check list (counts feed the inlier-container bookkeeping, so they must be exact) re-runs the
		// group synchronously: half the seats at a time, or (one seat left) half the samples -- a sample is a function of
		// its index, so batch boundaries change no result.
		void rerun_batch(std::vector<Request *> &g, int B)
		{
			if (g.size() > 1)
			{
				std::vector<Request *> a(g.begin(), g.begin() + g.size() / 2), b(g.begin() + g.size() / 2, g.end());
				rerun_batch_part(a, B);
				rerun_batch_part(b, B);
			}
			else if (B > 1)
				rerun_batch_part(g, (B + 1) / 2);
			else
			{
				error = "gcr_score: the FP64 re-check list overflows for a single sample";
				complete(g, false);
			}
		}
		void rerun_batch_part(std::vector<Request *> &g, int B)
		{
			Issued rec;
			rec.group = g;
			acquiring = Request::BATCH;
			rec.slot = &acquire();
			bool ok = do_batch(g, B, rec) && (gcr_sync(ctx) == GCR_OK || fail("gcr_sync"));
			bool completed = false;
			if (ok && rec.post)
				ok = rec.post(completed);
			if (!completed)
				complete(g, ok);
			rec.slot->busy = false;
		}
		bool do_batch(std::vector<Request *> &g, int B, Issued &rec)
		{
			Slot &sl = *rec.slot;
			InBlock &in = sl.in;
			const BatchSpec &s0 = *g[0]->bspec;
			if (s0.variant != 0)
			{
				error = "batch engine: the plane-and-parallax inner run works on a context of its own";
				return false;
			}
			const int m = s0.m, mm = s0.mm;
			const size_t K = (size_t)B * mm, srow = (size_t)B * m;
			const bool explicit_samples = s0.explicit_samples != nullptr;
			BatchClass *c = batch_class(B, m, mm);
			if (!c)
				return fail("gcr_malloc");
			if (seat_class.empty())
				seat_class.assign(S, 0);
			// inputs: per-seat first index / seed offset / active flag (+ the explicit index sets of samplers 2 and 4)
			in.reset();
			const size_t o_first = in.carve((size_t)S * 8), o_seed = in.carve((size_t)S * 8), o_active = in.carve((size_t)S);
			const size_t o_params_end = in.used;
			const size_t o_samples = explicit_samples ? in.carve((size_t)S * srow * sizeof(uint32_t)) : 0;
			if (!in.commit(ctx))
				return fail("pinned / device allocation");
			memset(in.host<uint8_t>(o_active), 0, (size_t)S);
			memset(in.host<uint64_t>(o_first), 0, (size_t)S * 8);
			memset(in.host<uint64_t>(o_seed), 0, (size_t)S * 8);
			int lo_seat = S, hi_seat = -1;
			for (Request *r : g)
			{
				const BatchSpec &s = *r->bspec;
				in.host<uint8_t>(o_active)[r->seat] = 1;
				// PROSAC draws are numbered from k = 1 (prosac_sampler.h:207-236); draw g of the run is PROSAC sample g + 1
				in.host<uint64_t>(o_first)[r->seat] = s.sampler == 1 ? s.first + 1 : s.first;
				// the kernels add seat * stride to the seed: pair p's streams are those of ITS seed, whatever seat it sits on
				in.host<uint64_t>(o_seed)[r->seat] = s.seed - (uint64_t)r->seat * kPairSeedStride;
				if (explicit_samples)
					memcpy(in.host<uint32_t>(o_samples) + (size_t)r->seat * srow, s.explicit_samples, srow * sizeof(uint32_t));
				lo_seat = std::min(lo_seat, r->seat);
				hi_seat = std::max(hi_seat, r->seat);
			}
			const size_t span = (size_t)(hi_seat - lo_seat + 1);
			if (!in.upload(ctx, 0, o_params_end))
				return fail("H2D");
			if (explicit_samples)
			{
				// the rows of the requesting seats (one copy over their span) straight into the class's sample array
				const size_t from = (size_t)lo_seat * srow * sizeof(uint32_t), bytes = span * srow * sizeof(uint32_t);
				if (gcr_memcpy_pinned_async(ctx, c->samples.as<char>() + from, in.host<char>(o_samples) + from, bytes, 1) != GCR_OK)
					return fail("H2D");
			}
			gcr_set_active_pairs(ctx, in.dev<uint8_t>(o_active));
			gcr_set_pair_streams(ctx, in.dev<uint64_t>(o_first), in.dev<uint64_t>(o_seed));
			bool ok = true;
			if (!explicit_samples)
				ok = gcr_sample(ctx, s0.sampler == 1 ? GCR_PROSAC : GCR_UNIFORM, m, 0, B, 0, c->samples.as<uint32_t>()) == GCR_OK;
			ok = ok && gcr_solve_minimal(ctx, problem, c->samples.as<uint32_t>(), B, c->models.as<double>(), c->valid.as<uint8_t>()) == GCR_OK;
			ok = ok && gcr_score(ctx, problem, s0.scoring, c->models.as<double>(), c->valid.as<uint8_t>(), (int)K, s0.threshold,
			                     GCR_SCORE_EXACT_COUNTS, c->scores.as<float>(), c->counts.as<uint32_t>()) == GCR_OK;
			gcr_set_active_pairs(ctx, nullptr);
			gcr_set_pair_streams(ctx, nullptr, nullptr);
			if (!ok)
				return fail("batch launch");
			// outputs: flags, FP32 scores, exact counts (+ the samples for DEGENSAC) of the seats [lo_seat, hi_seat], one copy each
			const bool want_samples = s0.want_samples && !explicit_samples;
			auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
			const size_t o_valid = 0, o_scores = pad(span * K), o_counts = o_scores + pad(span * K * 4), o_smp = o_counts + pad(span * K * 4),
			             o_dropped = o_smp + (want_samples ? pad(span * srow * 4) : 0);
			Staging &st = sl.st;
			if (!st.ensure(ctx, o_dropped + 256))
				return fail("pinned allocation");
			uint32_t *dropped = st.at<uint32_t>(o_dropped);
			*dropped = 0;
			if (gcr_memcpy_pinned_async(ctx, st.at<char>(o_valid), c->valid.as<char>() + (size_t)lo_seat * K, span * K, 0) != GCR_OK ||
			    gcr_memcpy_pinned_async(ctx, st.at<char>(o_scores), c->scores.as<char>() + (size_t)lo_seat * K * 4, span * K * 4, 0) != GCR_OK ||
			    gcr_memcpy_pinned_async(ctx, st.at<char>(o_counts), c->counts.as<char>() + (size_t)lo_seat * K * 4, span * K * 4, 0) != GCR_OK ||
			    (want_samples && gcr_memcpy_pinned_async(ctx, st.at<char>(o_smp), c->samples.as<char>() + (size_t)lo_seat * srow * 4, span * srow * 4, 0) != GCR_OK) ||
			    gcr_score_dropped_async(ctx, dropped) != GCR_OK)
				return fail("D2H");
			std::vector<Request *> group = g;
			Staging *stp = &st;
			rec.post = [=](bool &completed) mutable {
				if (*dropped)
				{
					completed = true;
					rerun_batch(group, B);
					return true;
				}
				for (Request *r : group)
				{
					seat_class[r->seat] = B;
					attach(r, *stp);
					BatchOut *out = r->bout;
					const size_t row = (size_t)(r->seat - lo_seat);
					const int ml_ = ml;
					r->deliver = [=]() {
						const uint8_t *v = stp->at<uint8_t>(o_valid) + row * K;
						const float *sc = stp->at<float>(o_scores) + row * K;
						const uint32_t *ct = stp->at<uint32_t>(o_counts) + row * K;
						out->valid.assign(v, v + K);
						out->scores.assign(sc, sc + K);
						out->counts.assign(ct, ct + K);
						out->models.resize(K * ml_);
						out->model_have.assign(K, 0); // the models stay on the device: fetched on demand (FETCH)
						if (want_samples)
						{
							const uint32_t *sp = stp->at<uint32_t>(o_smp) + row * srow;
							out->samples.assign(sp, sp + srow);
						}
						out->B = B;
					};
				}
				return true;
			};
			return true;
		}

		// ---- FETCH: model slots of the seats' current batches (one gather + one copy) ----
		bool do_fetch(std::vector<Request *> &g, Slot &sl)
		{
			InBlock &in = sl.in;
			const int B = seat_class.empty() ? 0 : seat_class[g[0]->seat];
			auto it = classes.find(B);
			if (B == 0 || it == classes.end())
			{
				error = "batch engine: fetch_models before any batch";
				return false;
			}
			BatchClass *c = it->second.get();
			size_t total = 0;
			for (Request *r : g)
				for (size_t slot : *r->slots)
					if (!r->bout->model_have[slot])
						++total;
			if (total == 0)
				return true;
			const size_t K = g[0]->bout->valid.size(); // slots per seat row of the seats' current batch
			in.reset();
			const size_t o_rows = in.carve(total * 8);
			if (!in.commit(ctx) || !sl.d_a.ensure(ctx, total * ml * sizeof(double)) || !sl.st.ensure(ctx, total * ml * sizeof(double)))
				return fail("allocation");
			size_t i = 0;
			for (Request *r : g)
				for (size_t slot : *r->slots)
					if (!r->bout->model_have[slot])
						in.host<uint64_t>(o_rows)[i++] = (uint64_t)r->seat * K + slot;
			if (!in.upload(ctx, 0, in.used) ||
			    gcr_gather_rows(ctx, c->models.p, (uint32_t)(ml * sizeof(double)), in.dev<uint64_t>(o_rows), (uint32_t)total, sl.d_a.p) != GCR_OK ||
			    gcr_memcpy_pinned_async(ctx, sl.st.p, sl.d_a.p, total * ml * sizeof(double), 0) != GCR_OK)
				return fail("fetch");
			i = 0;
			for (Request *r : g)
			{
				const size_t first = i;
				for (size_t slot : *r->slots)
					if (!r->bout->model_have[slot])
						++i;
				if (i == first)
					continue;
				attach(r, sl.st);
				BatchOut *out = r->bout;
				const std::vector<size_t> *slots_ = r->slots;
				Staging *st = &sl.st;
				const int ml_ = ml;
				r->deliver = [=]() {
					size_t j = first;
					for (size_t slot : *slots_)
						if (!out->model_have[slot])
						{
							memcpy(&out->models[slot * ml_], st->at<double>(0) + j * ml_, ml_ * sizeof(double));
							out->model_have[slot] = 1;
							++j;
						}
				};
			}
			return true;
		}

		// ---- EXACT: FP64 getScore of every requested model against its seat's points ----
		bool do_exact(std::vector<Request *> &g, Slot &sl)
		{
			InBlock &in = sl.in;
			size_t k = 0;
			bool any_idx = false, any_res = false;
			for (Request *r : g)
			{
				k += r->C;
				any_idx |= r->o_idx != nullptr;
				any_res |= r->o_res != nullptr;
			}
			in.reset();
			const size_t o_pairs = in.carve(k * sizeof(int32_t)), o_models = in.carve(k * ml * sizeof(double));
			if (!in.commit(ctx) || !sl.d_a.ensure(ctx, k * sizeof(double)) || !sl.d_b.ensure(ctx, k * sizeof(uint32_t)) ||
			    (any_idx && !sl.d_c.ensure(ctx, k * (size_t)n * sizeof(uint32_t))) || (any_res && !sl.d_d.ensure(ctx, k * (size_t)n * sizeof(double))))
				return fail("allocation");
			size_t i = 0;
			for (Request *r : g)
				for (size_t c = 0; c < r->C; ++c, ++i)
				{
					in.host<int32_t>(o_pairs)[i] = r->seat;
					memcpy(in.host<double>(o_models) + i * ml, r->models + c * ml, ml * sizeof(double));
				}
			const size_t o_s = 0, o_c = (k * 8 + 255) & ~(size_t)255, o_i = o_c + ((k * 4 + 255) & ~(size_t)255),
			             o_r = o_i + (any_idx ? ((k * (size_t)n * 4 + 255) & ~(size_t)255) : 0),
			             total = o_r + (any_res ? k * (size_t)n * 8 : 0) + 256;
			Staging &st_exact = sl.st;
			if (!st_exact.ensure(ctx, total))
				return fail("pinned allocation");
			if (!in.upload(ctx, 0, in.used) ||
			    gcr_score_exact_multi(ctx, problem, g[0]->scoring, in.dev<int32_t>(o_pairs), in.dev<double>(o_models), (int)k, g[0]->threshold,
			                          sl.d_a.as<double>(), sl.d_b.as<uint32_t>(), any_idx ? sl.d_c.as<uint32_t>() : nullptr,
			                          any_res ? sl.d_d.as<double>() : nullptr) != GCR_OK)
				return fail("gcr_score_exact_multi");
			bool ok = gcr_memcpy_pinned_async(ctx, st_exact.at<char>(o_s), sl.d_a.p, k * 8, 0) == GCR_OK &&
			          gcr_memcpy_pinned_async(ctx, st_exact.at<char>(o_c), sl.d_b.p, k * 4, 0) == GCR_OK;
			if (ok && any_idx)
				ok = gcr_memcpy_pinned_async(ctx, st_exact.at<char>(o_i), sl.d_c.p, k * (size_t)n * 4, 0) == GCR_OK;
			i = 0;
			for (Request *r : g) // residual rows: only the few requests that asked (IRLS weights, the post-run polish)
			{
				if (ok && r->o_res)
					ok = gcr_memcpy_pinned_async(ctx, st_exact.at<char>(o_r) + i * (size_t)n * 8, sl.d_d.as<char>() + i * (size_t)n * 8, (size_t)n * 8, 0) == GCR_OK;
				i += r->C;
			}
			if (!ok)
				return fail("D2H");
			i = 0;
			for (Request *r : g)
			{
				attach(r, st_exact);
				Staging *st = &st_exact;
				const size_t first = i, C = r->C, nn = (size_t)n;
				double *o_score = r->o_score, *o_res = r->o_res;
				uint32_t *o_count = r->o_count, *o_idx = r->o_idx;
				r->deliver = [=]() {
					for (size_t c = 0; c < C; ++c)
					{
						const uint32_t cnt = st->at<uint32_t>(o_c)[first + c];
						o_score[c] = st->at<double>(o_s)[first + c];
						o_count[c] = cnt;
						if (o_idx)
							memcpy(o_idx + c * nn, st->at<uint32_t>(o_i) + (first + c) * nn, (size_t)cnt * sizeof(uint32_t));
					}
					if (o_res)
						memcpy(o_res, st->at<double>(o_r) + first * nn, nn * sizeof(double));
				};
				i += C;
			}
			return true;
		}

		// ---- LABEL: labeling()'s unary terms + clamped distances per requested model ----
		bool do_label(std::vector<Request *> &g, Slot &sl)
		{
			InBlock &in = sl.in;
			const size_t k = g.size(), row = 3 * (size_t)n;
			in.reset();
			const size_t o_pairs = in.carve(k * sizeof(int32_t)), o_models = in.carve(k * ml * sizeof(double));
			if (!in.commit(ctx) || !sl.d_a.ensure(ctx, k * row * sizeof(double)) || !sl.st.ensure(ctx, k * row * sizeof(double)))
				return fail("allocation");
			for (size_t i = 0; i < k; ++i)
			{
				in.host<int32_t>(o_pairs)[i] = g[i]->seat;
				memcpy(in.host<double>(o_models) + i * ml, g[i]->models, ml * sizeof(double));
			}
			if (!in.upload(ctx, 0, in.used) ||
			    gcr_labeling_terms_multi(ctx, problem, in.dev<int32_t>(o_pairs), in.dev<double>(o_models), (int)k, g[0]->threshold, g[0]->lambda,
			                             sl.d_a.as<double>()) != GCR_OK ||
			    gcr_memcpy_pinned_async(ctx, sl.st.p, sl.d_a.p, k * row * sizeof(double), 0) != GCR_OK)
				return fail("gcr_labeling_terms_multi");
			for (size_t i = 0; i < k; ++i)
			{
				attach(g[i], sl.st);
				Staging *st = &sl.st;
				std::vector<double> *out = g[i]->o_unary;
				g[i]->deliver = [=]() { out->assign(st->at<double>(0) + i * row, st->at<double>(0) + (i + 1) * row); };
			}
			return true;
		}

		// ---- REFIT: non-minimal fits of the requests' index sets ----
		bool do_refit(std::vector<Request *> &g, Slot &sl)
		{
			InBlock &in = sl.in;
			const Request &r0 = *g[0];
			const int rm = gcr_max_refit_models(r0.rp), rml = gcr_model_len(r0.rp);
			const bool has_w = r0.weights != nullptr;
			const int wide = r0.sizes->size() == 1; // the CTA shape of a one-set call of the unbatched driver
			size_t n_sets = 0, n_idx = 0, n_w = 0;
			for (Request *r : g)
			{
				n_sets += r->sizes->size();
				size_t end = 0;
				for (size_t s = 0; s < r->sizes->size(); ++s)
					end = std::max(end, (size_t)(*r->offsets)[s] + (*r->sizes)[s]);
				n_idx += end;
				n_w += r->n_weights;
			}
			in.reset();
			const size_t o_pair = in.carve(n_sets * 4), o_off = in.carve(n_sets * 4), o_size = in.carve(n_sets * 4), o_woff = in.carve(n_sets * 4),
			             o_idx = in.carve((n_idx ? n_idx : 1) * 4), o_w = in.carve((n_w ? n_w : 1) * 8);
			const size_t o_ok = (n_sets * rm * rml * sizeof(double) + 255) & ~(size_t)255;
			if (!in.commit(ctx) || !sl.d_a.ensure(ctx, n_sets * rm * rml * sizeof(double)) || !sl.d_b.ensure(ctx, n_sets * rm) ||
			    !sl.st.ensure(ctx, o_ok + n_sets * rm + 256))
				return fail("allocation");
			size_t s_at = 0, i_at = 0, w_at = 0;
			for (Request *r : g)
			{
				size_t end = 0;
				for (size_t s = 0; s < r->sizes->size(); ++s)
				{
					in.host<int32_t>(o_pair)[s_at] = r->seat;
					in.host<uint32_t>(o_off)[s_at] = (uint32_t)(i_at + (*r->offsets)[s]);
					in.host<uint32_t>(o_size)[s_at] = (*r->sizes)[s];
					in.host<uint32_t>(o_woff)[s_at] = (uint32_t)w_at;
					end = std::max(end, (size_t)(*r->offsets)[s] + (*r->sizes)[s]);
					++s_at;
				}
				memcpy(in.host<uint32_t>(o_idx) + i_at, r->flat, end * sizeof(uint32_t));
				i_at += end;
				if (has_w)
				{
					memcpy(in.host<double>(o_w) + w_at, r->weights, r->n_weights * sizeof(double));
					w_at += r->n_weights;
				}
			}
			if (!in.upload(ctx, 0, in.used) ||
			    gcr_refit_multi(ctx, r0.rp, r0.ba_frame, in.dev<int32_t>(o_pair), in.dev<uint32_t>(o_idx), in.dev<uint32_t>(o_off),
			                    in.dev<uint32_t>(o_size), (int)n_sets, has_w ? in.dev<double>(o_w) : nullptr, has_w ? in.dev<uint32_t>(o_woff) : nullptr,
			                    has_w ? r0.weight_mode : (int)GCR_WEIGHT_NONE, nullptr, wide, sl.d_a.as<double>(), sl.d_b.as<uint8_t>()) != GCR_OK ||
			    gcr_memcpy_pinned_async(ctx, sl.st.p, sl.d_a.p, n_sets * rm * rml * sizeof(double), 0) != GCR_OK ||
			    gcr_memcpy_pinned_async(ctx, sl.st.at<char>(o_ok), sl.d_b.p, n_sets * rm, 0) != GCR_OK)
				return fail("gcr_refit_multi");
			s_at = 0;
			for (Request *r : g)
			{
				attach(r, sl.st);
				Staging *st = &sl.st;
				const size_t first = s_at, cnt = r->sizes->size();
				std::vector<double> *om = r->o_models;
				std::vector<uint8_t> *ook = r->o_ok;
				r->deliver = [=]() {
					om->assign(st->at<double>(0) + first * rm * rml, st->at<double>(0) + (first + cnt) * rm * rml);
					ook->assign(st->at<uint8_t>(o_ok) + first * rm, st->at<uint8_t>(o_ok) + (first + cnt) * rm);
				};
				s_at += cnt;
			}
			return true;
		}

		// ---- LO: the inner RANSAC of graphCutLocalOptimization for every requesting seat: pool samples -> refits -> FP32 scores ----
		// (the model rows live in a per-engine [S][trials rm] array: gcr_score works on every pair of the context)
		DeviceBuffer lo_models, lo_ok, lo_scores;
		bool do_lo(std::vector<Request *> &g, Slot &sl)
		{
			InBlock &in = sl.in;
			const LoSpec &s0 = *g[0]->lo;
			const bool pool = s0.sample_pool;
			const int trials = pool ? s0.trials : 1, rm = s0.rm, rml = gcr_model_len(s0.refit_problem);
			const size_t nreq = g.size(), n_sets = nreq * (size_t)trials, per_seat = (size_t)trials * rm;
			size_t n_pool = 0, n_draws = 0;
			int lo_seat = S, hi_seat = -1;
			for (Request *r : g)
			{
				n_pool += r->lo->pool_size;
				n_draws += (size_t)trials * r->lo->sample_size;
				lo_seat = std::min(lo_seat, r->seat);
				hi_seat = std::max(hi_seat, r->seat);
			}
			const size_t span = (size_t)(hi_seat - lo_seat + 1);
			in.reset();
			const size_t o_pools = in.carve((n_pool ? n_pool : 1) * 4), o_reqs = in.carve(nreq * sizeof(gcr_pool_request)), o_pair = in.carve(n_sets * 4),
			             o_off = in.carve(n_sets * 4), o_size = in.carve(n_sets * 4), o_slot = in.carve(n_sets * 4);
			const size_t row_m = per_seat * rml * sizeof(double), row_ok = per_seat, row_sc = per_seat * sizeof(float);
			const size_t o_m = 0, o_k = (span * row_m + 255) & ~(size_t)255, o_s = o_k + ((span * row_ok + 255) & ~(size_t)255);
			if (!in.commit(ctx) || !lo_models.ensure(ctx, (size_t)S * row_m) || !lo_ok.ensure(ctx, (size_t)S * row_ok) ||
			    !lo_scores.ensure(ctx, (size_t)S * row_sc) || (pool && !sl.d_d.ensure(ctx, (n_draws ? n_draws : 1) * 4)) ||
			    !sl.st.ensure(ctx, o_s + span * row_sc + 256))
				return fail("allocation");
			size_t p_at = 0, d_at = 0, s_at = 0;
			for (size_t q = 0; q < nreq; ++q)
			{
				const Request *r = g[q];
				const LoSpec &s = *r->lo;
				memcpy(in.host<uint32_t>(o_pools) + p_at, s.pool, s.pool_size * sizeof(uint32_t));
				gcr_pool_request &pr = in.host<gcr_pool_request>(o_reqs)[q];
				pr.pool_offset = (uint32_t)p_at;
				pr.pool_size = (uint32_t)s.pool_size;
				pr.sample_size = (uint32_t)s.sample_size;
				pr.out_offset = (uint32_t)d_at;
				pr.first_index = s.first_index;
				pr.seed = s.seed;
				pr.stream = s.stream;
				pr.reserved = 0;
				for (int t = 0; t < trials; ++t, ++s_at)
				{
					in.host<int32_t>(o_pair)[s_at] = r->seat;
					// pool samples: set t of the request in the draw buffer; whole-pool fit: the pool itself
					in.host<uint32_t>(o_off)[s_at] = pool ? (uint32_t)(d_at + (size_t)t * s.sample_size) : (uint32_t)p_at;
					in.host<uint32_t>(o_size)[s_at] = pool ? (uint32_t)s.sample_size : (uint32_t)s.pool_size;
					in.host<uint32_t>(o_slot)[s_at] = (uint32_t)((size_t)r->seat * trials + t);
				}
				p_at += s.pool_size;
				d_at += (size_t)trials * s.sample_size;
			}
			// a fit of ONE set per call takes the wide CTA shape in the unbatched driver, many sets one warp each (the block sums
			// associate by shape)
			const int wide = trials == 1;
			bool ok = in.upload(ctx, 0, in.used) && gcr_memset(ctx, lo_ok.p, 0, (size_t)S * row_ok) == GCR_OK;
			if (ok && pool)
				ok = gcr_sample_from_pools(ctx, in.dev<uint32_t>(o_pools), in.dev<gcr_pool_request>(o_reqs), (int)nreq, trials, sl.d_d.as<uint32_t>()) == GCR_OK;
			ok = ok && gcr_refit_multi(ctx, s0.refit_problem, s0.ba_frame, in.dev<int32_t>(o_pair), pool ? sl.d_d.as<uint32_t>() : in.dev<uint32_t>(o_pools),
			                           in.dev<uint32_t>(o_off), in.dev<uint32_t>(o_size), (int)n_sets, nullptr, nullptr, GCR_WEIGHT_NONE,
			                           in.dev<uint32_t>(o_slot), wide, lo_models.as<double>(), lo_ok.as<uint8_t>()) == GCR_OK;
			ok = ok && gcr_score(ctx, problem, s0.scoring, lo_models.as<double>(), lo_ok.as<uint8_t>(), (int)per_seat, s0.threshold, 0,
			                     lo_scores.as<float>(), nullptr) == GCR_OK;
			ok = ok && gcr_memcpy_pinned_async(ctx, sl.st.at<char>(o_m), lo_models.as<char>() + (size_t)lo_seat * row_m, span * row_m, 0) == GCR_OK &&
			     gcr_memcpy_pinned_async(ctx, sl.st.at<char>(o_k), lo_ok.as<char>() + (size_t)lo_seat * row_ok, span * row_ok, 0) == GCR_OK &&
			     gcr_memcpy_pinned_async(ctx, sl.st.at<char>(o_s), lo_scores.as<char>() + (size_t)lo_seat * row_sc, span * row_sc, 0) == GCR_OK;
			if (!ok)
				return fail("local optimisation batch");
			for (Request *r : g)
			{
				attach(r, sl.st);
				Staging *st = &sl.st;
				const size_t row = (size_t)(r->seat - lo_seat);
				std::vector<double> *om = r->o_models;
				std::vector<uint8_t> *ook = r->o_ok;
				std::vector<float> *osc = r->o_scores32;
				r->deliver = [=]() {
					const double *m = st->at<double>(o_m) + row * per_seat * rml;
					const uint8_t *k = st->at<uint8_t>(o_k) + row * per_seat;
					const float *sc = st->at<float>(o_s) + row * per_seat;
					om->assign(m, m + per_seat * rml);
					ook->assign(k, k + per_seat);
					osc->assign(sc, sc + per_seat);
				};
			}
			return true;
		}
	};

	// ------------------------------------------------------------------------------------------------ backend ----
	class BatchedBackend : public Backend
	{
	public:
		Engine *E;
		int seat;
		BatchedBackend(Engine *e, int seat_) : E(e), seat(seat_) { keeps_edges = false; }

		bool post(Request &r)
		{
			r.seat = seat;
			if (E->submit(r))
				return true;
			error = r.error;
			return false;
		}
		bool serial(std::function<bool(DirectBackend &)> fn)
		{
			Request r;
			r.kind = Request::SERIAL;
			r.fn = std::move(fn);
			return post(r);
		}
		bool upload(const double *pts, int, int) override
		{
			Request r;
			r.kind = Request::UPLOAD;
			r.pts = pts;
			return post(r);
		}
		bool set_point_probabilities(const double *, int) override
		{
			error = "sampler 3 (one probability table per context) is not available in a batch";
			return false;
		}
		bool grid_device(const double cell_sizes[4], uint64_t cells, std::vector<uint32_t> &sorted_idx, std::vector<uint32_t> &cell_offsets,
		                 uint32_t &n_cells, uint64_t &neighbor_number, std::vector<uint32_t> *edges) override
		{
			return serial([&](DirectBackend &d) { return d.grid_device(cell_sizes, cells, sorted_idx, cell_offsets, n_cells, neighbor_number, edges); });
		}
		bool radius_graph(double radius, std::vector<uint32_t> &offsets, std::vector<uint32_t> &neighbors, std::vector<uint32_t> &edges) override
		{
			return serial([&](DirectBackend &d) { return d.radius_graph(radius, offsets, neighbors, edges); });
		}
		bool set_edges(const std::vector<uint32_t> &) override { return true; } // the host keeps the edge list (keeps_edges = false)
		bool run_batch(const BatchSpec &spec, BatchOut &out) override
		{
			Request r;
			r.kind = Request::BATCH;
			r.bspec = &spec;
			r.bout = &out;
			return post(r);
		}
		bool fetch_models(const std::vector<size_t> &slots, BatchOut &out) override
		{
			bool need = false;
			for (size_t s : slots)
				need |= !out.model_have[s];
			if (!need)
				return true;
			Request r;
			r.kind = Request::FETCH;
			r.slots = &slots;
			r.bout = &out;
			return post(r);
		}
		bool exact(const double *model, int scoring, double, double threshold, double &score, uint32_t &count, uint32_t *inliers,
		           double *residuals) override
		{
			Request r;
			r.kind = Request::EXACT;
			r.models = model;
			r.C = 1;
			r.scoring = scoring;
			r.threshold = threshold;
			r.o_score = &score;
			r.o_count = &count;
			r.o_idx = inliers;
			r.o_res = residuals;
			return post(r);
		}
		bool exact_many(const double *models, size_t C, int scoring, double threshold, std::vector<double> &score, std::vector<uint32_t> &count,
		                std::vector<uint32_t> &idx) override
		{
			score.resize(C);
			count.resize(C);
			idx.resize(C * (size_t)n);
			Request r;
			r.kind = Request::EXACT;
			r.models = models;
			r.C = C;
			r.scoring = scoring;
			r.threshold = threshold;
			r.o_score = score.data();
			r.o_count = count.data();
			r.o_idx = idx.data();
			return post(r);
		}
		bool labeling_terms(const double *model, double threshold, double lambda, size_t, bool, std::vector<double> &unary3n,
		                    std::vector<double> &pairwise) override
		{
			pairwise.clear();
			Request r;
			r.kind = Request::LABEL;
			r.models = model;
			r.threshold = threshold;
			r.lambda = lambda;
			r.o_unary = &unary3n;
			return post(r);
		}
		bool refit_sets(int rp, int ba_frame, const uint32_t *flat, const std::vector<uint32_t> &offsets, const std::vector<uint32_t> &sizes,
		                const double *weights, size_t n_weights, int weight_mode, std::vector<double> &models, std::vector<uint8_t> &ok) override
		{
			if (rp == GCR_PNP)
			{
				error = "the absolute-pose problem is not available in a batch";
				return false;
			}
			Request r;
			r.kind = Request::REFIT;
			r.rp = rp;
			r.ba_frame = ba_frame;
			r.flat = flat;
			r.offsets = &offsets;
			r.sizes = &sizes;
			r.weights = weights;
			r.n_weights = weights ? n_weights : 0;
			r.weight_mode = weight_mode;
			r.o_models = &models;
			r.o_ok = &ok;
			return post(r);
		}
		bool lo_batch(const LoSpec &spec, std::vector<double> &models, std::vector<uint8_t> &ok, std::vector<float> &scores32) override
		{
			Request r;
			r.kind = Request::LO;
			r.lo = &spec;
			r.o_models = &models;
			r.o_ok = &ok;
			r.o_scores32 = &scores32;
			return post(r);
		}
		bool refit_e_ba(const uint32_t *set, size_t count, const double *weights, std::vector<double> &models, std::vector<uint8_t> &ok) override
		{
			return serial([&](DirectBackend &d) { return d.refit_e_ba(set, count, weights, models, ok); });
		}
	};

	// a worker's private single-pair context: DEGENSAC's inner run (its plane-and-parallax solver is not a batched operation).
	// The contexts live in a process-wide pool: creating one allocates pinned and device memory, which stalls every stream.
	std::mutex g_nested_mu;
	std::vector<std::pair<int, gcr_ctx *>> g_nested_free; // (device, context)
	struct WorkerState
	{
		int device = 0;
		gcr_ctx *ctx = nullptr;
		std::unique_ptr<DirectBackend> be;
		bool acquire()
		{
			{
				std::lock_guard<std::mutex> lock(g_nested_mu);
				for (size_t i = 0; i < g_nested_free.size(); ++i)
					if (g_nested_free[i].first == device)
					{
						ctx = g_nested_free[i].second;
						g_nested_free.erase(g_nested_free.begin() + i);
						return true;
					}
			}
			return gcr_ctx_create(device, &ctx) == GCR_OK;
		}
		~WorkerState()
		{
			be.reset(); // (its device blocks go back to the context's cache)
			if (ctx)
			{
				std::lock_guard<std::mutex> lock(g_nested_mu);
				g_nested_free.push_back({device, ctx});
			}
		}
	};
	Backend *nested_backend_of(Driver *d)
	{
		WorkerState *ws = static_cast<WorkerState *>(d->owner);
		if (!ws->ctx)
		{
			if (!ws->acquire())
			{
				d->error = "gcr_ctx_create failed (DEGENSAC's inner run)";
				return nullptr;
			}
			ws->be.reset(new DirectBackend(ws->ctx));
		}
		ws->be->problem = GCR_F;
		ws->be->n = d->n;
		ws->be->ml = 9;
		ws->be->cols = 4;
		if (!ws->be->upload(d->pts, d->n, 4))
		{
			d->error = ws->be->error;
			return nullptr;
		}
		return ws->be.get();
	}

	// engines persist between calls (contexts, class buffers, pinned staging: ~100 ms to set up, a call of 128 pairs is ~10 ms)
	struct EngineKey
	{
		int device, problem, n, seats, index;
		bool operator<(const EngineKey &o) const
		{
			return std::tie(device, problem, n, seats, index) < std::tie(o.device, o.problem, o.n, o.seats, o.index);
		}
	};
	std::mutex g_engine_mu, g_call_mu; // (one batch call at a time per process: the cached engines are shared)
	std::map<EngineKey, std::unique_ptr<Engine>> g_engines;

	void inverse3(const double *M, double *R)
	{
		// Eigen's 3 x 3 inverse: cofactors / determinant
		const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
		const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
		const double inv = 1.0 / det;
		R[0] = c00 * inv; R[1] = (M[2] * M[7] - M[1] * M[8]) * inv; R[2] = (M[1] * M[5] - M[2] * M[4]) * inv;
		R[3] = c01 * inv; R[4] = (M[0] * M[8] - M[2] * M[6]) * inv; R[5] = (M[2] * M[3] - M[0] * M[5]) * inv;
		R[6] = c02 * inv; R[7] = (M[1] * M[6] - M[0] * M[7]) * inv; R[8] = (M[0] * M[4] - M[1] * M[3]) * inv;
	}

	int find_batch(int device, int problem, const double *h_points, int n, int n_pairs, const double *K1, const double *K2, int k_stride,
	               const gcr_run_settings *settings, const uint64_t *seeds, int seats, int engines, double *models_out, uint8_t *inlier_masks,
	               int *inlier_counts, gcr_run_stats *stats)
	{
		if (!h_points || !settings || !models_out || !inlier_masks || !inlier_counts || n <= 0 || n_pairs <= 0)
			return -2;
		if (problem != GCR_H && problem != GCR_F && problem != GCR_E)
			return -3;
		if (settings->sampler == 3)
		{
			g_driver_error = "gcr_find_model_batch: sampler 3 keeps one probability table per context; use gcr_find_model";
			return -3;
		}
		if (n > 32768)
		{
			g_driver_error = "gcr_find_model_batch: at most 32768 correspondences per pair (gcr_points_replace_pairs)";
			return -3;
		}
		if (engines <= 0)
		{
			const char *e = getenv("GCR_BATCH_ENGINES");
			engines = e ? atoi(e) : 2;
		}
		if (seats <= 0)
		{
			const char *e = getenv("GCR_BATCH_SEATS");
			seats = e ? atoi(e) : 512;
		}
		std::lock_guard<std::mutex> call_lock(g_call_mu);
		const auto call_t0 = std::chrono::steady_clock::now();
		engines = std::max(1, std::min(engines, n_pairs));
		int per_engine = std::max(1, std::min(seats, n_pairs) / engines);
		const size_t row = (size_t)n * 4;
		std::vector<Engine *> eng(engines);
		{
			std::lock_guard<std::mutex> lock(g_engine_mu);
			for (int e = 0; e < engines; ++e)
			{
				const EngineKey key{device, problem, n, per_engine, e};
				std::unique_ptr<Engine> &slot = g_engines[key];
				if (!slot)
				{
					slot.reset(new Engine());
					if (!slot->create(device, problem, n, per_engine, h_points + (size_t)(e * per_engine) * row))
					{
						g_driver_error = slot->error;
						slot.reset();
						g_engines.erase(key);
						return -1;
					}
				}
				eng[e] = slot.get();
			}
		}
		std::atomic<int> next_pair{0};
		std::atomic<bool> abort{false};
		std::mutex err_mu;
		std::string first_error;
		auto worker = [&](Engine *E, int seat) {
			BatchedBackend be(E, seat);
			WorkerState ws;
			ws.device = device;
			std::vector<double> normalized;
			for (;;)
			{
				const int p = next_pair.fetch_add(1);
				if (p >= n_pairs || abort.load())
					break;
				Driver d;
				d.be = &be;
				d.problem = problem;
				d.n = n;
				d.pts = h_points + (size_t)p * row;
				d.S = *settings;
				d.S.seed = seeds ? seeds[p] : settings->seed + (uint64_t)p * kPairSeedStride;
				d.nested_backend = nested_backend_of;
				d.owner = &ws;
				if (K1)
				{
					// findEssentialMatrix_: K^-1 normalisation (utils.h:65-110), threshold / mean focal length (py/gcransac_python.cpp:2624-2626)
					const double *k1 = K1 + (size_t)p * k_stride, *k2 = K2 + (size_t)p * k_stride;
					double K1i[9], K2i[9];
					inverse3(k1, K1i);
					inverse3(k2, K2i);
					normalized.resize(row);
					const double *src = d.pts;
					for (int r = 0; r < n; ++r)
					{
						const double x0 = src[4 * (size_t)r], y0 = src[4 * (size_t)r + 1], x1 = src[4 * (size_t)r + 2], y1 = src[4 * (size_t)r + 3];
						normalized[4 * (size_t)r + 0] = K1i[0] * x0 + K1i[1] * y0 + K1i[2] * 1.0;
						normalized[4 * (size_t)r + 1] = K1i[3] * x0 + K1i[4] * y0 + K1i[5] * 1.0;
						normalized[4 * (size_t)r + 2] = K2i[0] * x1 + K2i[1] * y1 + K2i[2] * 1.0;
						normalized[4 * (size_t)r + 3] = K2i[3] * x1 + K2i[4] * y1 + K2i[5] * 1.0;
					}
					d.grid_pts = src;
					d.pts = normalized.data();
					d.S.threshold = settings->threshold / (0.25 * (k1[0] + k1[4] + k2[0] + k2[4]));
				}
				const int r = d.run(models_out + (size_t)p * 9, inlier_masks + (size_t)p * n);
				inlier_counts[p] = r < 0 ? 0 : r;
				if (stats)
					stats[p] = d.st;
				if (r < 0)
				{
					std::lock_guard<std::mutex> lock(err_mu);
					if (first_error.empty())
						first_error = d.error;
					abort.store(true);
				}
			}
			E->worker_done();
		};
		// one fiber per seat, a few OS threads (about one per core, the dispatchers counted) run whichever are runnable
		RunQueue rq;
		std::vector<std::unique_ptr<Fiber>> fibers;
		for (int e = 0; e < engines; ++e)
			for (int s = 0; s < per_engine; ++s)
			{
				fibers.emplace_back(new Fiber());
				Fiber *f = fibers.back().get();
				Engine *E = eng[e];
				f->body = [&worker, E, s]() { worker(E, s); };
				if (!fiber_start(f, (size_t)1 << 20))
				{
					g_driver_error = "gcr_find_model_batch: no memory for the fiber stacks";
					return -1;
				}
			}
		rq.live = (int)fibers.size();
		for (auto &f : fibers)
			rq.q.push_back(f.get());
		int n_threads = getenv("GCR_BATCH_THREADS") ? atoi(getenv("GCR_BATCH_THREADS")) : (int)std::thread::hardware_concurrency() - engines;
		n_threads = std::max(1, std::min(n_threads, (int)fibers.size()));
		std::vector<std::thread> threads;
		for (int e = 0; e < engines; ++e)
		{
			eng[e]->run_queue = &rq;
			for (int s = 0; s < per_engine; ++s)
				eng[e]->worker_started();
			eng[e]->dispatcher = std::thread([E = eng[e]]() { E->run_dispatcher(); });
		}
		for (int t = 0; t < n_threads; ++t)
			threads.emplace_back(fiber_thread, &rq);
		const double spawn_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - call_t0).count();
		for (std::thread &t : threads)
			t.join();
		for (int e = 0; e < engines; ++e)
			eng[e]->dispatcher.join();
		if (getenv("GCR_BATCH_STATS"))
		{
			static const char *names[Request::KINDS] = {"upload", "batch", "exact", "label", "refit", "lo", "fetch", "serial"};
			fprintf(stderr, "batch call: %d pairs, %.2f ms (threads up after %.2f ms)\n", n_pairs,
			        1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - call_t0).count(), 1e3 * spawn_seconds);
			for (int e = 0; e < engines; ++e)
			{
				Engine *E = eng[e];
				fprintf(stderr, "  engine %d (cumulative): %llu super-steps, %llu groups, %llu requests, dispatcher idle %.1f ms;", e,
				        (unsigned long long)E->super_steps, (unsigned long long)E->groups, (unsigned long long)E->requests, 1e3 * E->idle_seconds);
				fprintf(stderr, " sync %.1f ms; complete %.1f ms;", 1e3 * E->sync_seconds, 1e3 * E->complete_seconds);
				for (int k = 0; k < Request::KINDS; ++k)
					if (E->kind_groups[k])
						fprintf(stderr, " %s %llu req / %llu grp %.1f ms;", names[k], (unsigned long long)E->kind_requests[k],
						        (unsigned long long)E->kind_groups[k], 1e3 * E->kind_seconds[k]);
				fprintf(stderr, "\n");
			}
		}
		if (!first_error.empty())
		{
			g_driver_error = first_error;
			fprintf(stderr, "gcr_find_model_batch: %s\n", first_error.c_str());
			return -1;
		}
		return 0;
	}
}

extern "C" int gcr_find_model_batch(int device, int problem, const double *h_points, int n, int n_pairs, const gcr_run_settings *settings,
                                    const uint64_t *seeds, int seats, int engines, double *models_out, uint8_t *inlier_masks, int *inlier_counts,
                                    gcr_run_stats *stats)
{
	return find_batch(device, problem, h_points, n, n_pairs, nullptr, nullptr, 0, settings, seeds, seats, engines, models_out, inlier_masks,
	                  inlier_counts, stats);
}

extern "C" int gcr_find_essential_batch(int device, const double *h_points, int n, int n_pairs, const double *K1, const double *K2, int per_pair_K,
                                        const gcr_run_settings *settings, const uint64_t *seeds, int seats, int engines, double *models_out,
                                        uint8_t *inlier_masks, int *inlier_counts, gcr_run_stats *stats)
{
	if (!K1 || !K2)
		return -2;
	return find_batch(device, GCR_E, h_points, n, n_pairs, K1, K2, per_pair_K ? 9 : 0, settings, seeds, seats, engines, models_out, inlier_masks,
	                  inlier_counts, stats);
}

extern "C" void gcr_batch_engines_release(void)
{
	std::lock_guard<std::mutex> lock(g_engine_mu);
	g_engines.clear();
	std::lock_guard<std::mutex> lock2(g_nested_mu);
	for (auto &c : g_nested_free)
		gcr_ctx_destroy(c.second);
	g_nested_free.clear();
}

// The engine's host-side concurrency core without a GPU: `n_fibers` fibers each post `rounds` requests (a token in a
// pending list) and switch out; a dispatcher thread completes whatever is pending; `n_threads` OS threads run the fibers.
// Same Fiber / RunQueue / state protocol as Engine::submit and Engine::complete.  Returns the number of completed requests
// (n_fibers x rounds when nothing was lost or resumed twice), < 0 on a set-up error.  For the CPU tier of the tests.
extern "C" long long gcr_fiber_selftest(int n_fibers, int n_threads, int rounds)
{
	if (n_fibers <= 0 || n_threads <= 0 || rounds < 0)
		return -1;
	RunQueue rq;
	std::mutex mu;
	std::condition_variable cv;
	std::vector<Fiber *> pending;
	int live = n_fibers;
	std::atomic<long long> total{0};
	std::vector<long long> per_fiber(n_fibers, 0);
	std::vector<std::unique_ptr<Fiber>> fibers;
	for (int i = 0; i < n_fibers; ++i)
	{
		fibers.emplace_back(new Fiber());
		Fiber *f = fibers.back().get();
		f->body = [&, i]() {
			for (int r = 0; r < rounds; ++r)
			{
				Fiber *self = current_fiber();
				{
					std::lock_guard<std::mutex> lock(mu);
					pending.push_back(self);
					cv.notify_one();
				}
				swapcontext(&self->uc, current_scheduler()); // back when the dispatcher has completed the request (on any thread)
				++per_fiber[i];
				total.fetch_add(1, std::memory_order_relaxed);
			}
			std::lock_guard<std::mutex> lock(mu);
			--live;
			cv.notify_one();
		};
		if (!fiber_start(f, (size_t)256 << 10))
			return -2;
	}
	rq.live = n_fibers;
	for (auto &f : fibers)
		rq.q.push_back(f.get());
	std::thread dispatcher([&]() {
		std::vector<Fiber *> batch;
		for (;;)
		{
			{
				std::unique_lock<std::mutex> lock(mu);
				cv.wait(lock, [&] { return !pending.empty() || live == 0; });
				if (pending.empty() && live == 0)
					return;
				batch.swap(pending);
			}
			for (Fiber *f : batch)
				if (f->state.exchange(Fiber::READY, std::memory_order_acq_rel) == Fiber::PARKED) // (else its thread enqueues it)
				{
					f->state.store(Fiber::RUNNING, std::memory_order_release);
					rq.push(f);
				}
			batch.clear();
		}
	});
	std::vector<std::thread> threads;
	for (int t = 0; t < n_threads; ++t)
		threads.emplace_back(fiber_thread, &rq);
	for (std::thread &t : threads)
		t.join();
	dispatcher.join();
	for (int i = 0; i < n_fibers; ++i)
		if (per_fiber[i] != rounds)
			return -3;
	return total.load();
}

extern "C" uint64_t gcr_batch_launch_count(void)
{
	std::lock_guard<std::mutex> lock(g_engine_mu);
	uint64_t total = 0;
	for (const auto &kv : g_engines)
		if (kv.second && kv.second->ctx)
			total += gcr_launch_count(kv.second->ctx);
	return total;
}

extern "C" const char *gcr_driver_last_error(void) { return g_driver_error.c_str(); }
#endif
