// pz_runtime.cu -- device, memory pool, copies, streams/events.
// B200-native replacement of the reference's Cuda/Source/Core/{Device,Buffer,Allocator,Stream}.c
// surface, exposed as plain C entry points (no CPython objects; the ctypes shim owns lifetimes).
#include "pz_common.h"

#include <cstdarg>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void pz_set_error(int code, const char* fmt, ...)
{
	(void)code;
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

static cudaStream_t g_default_stream = nullptr;

cudaStream_t pz_stream(void* s) { return s ? (cudaStream_t)s : g_default_stream; }

void pz_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int pz_num_sms()
{
	static int sms = 0;
	if (sms == 0) {
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess) return 148;
		if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
	}
	return sms;
}

// Library-owned scratch, one buffer per device.  Use is stream-ordered: the kernel that writes it and the kernel that reads it
// are enqueued back to back on the caller's stream (the reference runs everything on the legacy default stream), and the next
// user overwrites it only after both have run.  Growth never frees: a captured CUDA graph (pz_graph_*) keeps the address it
// saw, so an outgrown block is retired (kept until process exit) instead of being handed back to the driver -- and since
// cudaMalloc is not allowed while a stream is capturing, growth under capture is an error the caller sees (StepGraph warms the
// step up eagerly first, which sizes the scratch).
namespace {
struct DeviceScratch { void* buf = nullptr; size_t cap = 0; std::vector<void*> retired; };
DeviceScratch g_scratch[64];
std::mutex g_scratch_mu;
}

void* pz_scratch(size_t bytes)
{
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return nullptr; }
	std::lock_guard<std::mutex> lock(g_scratch_mu);
	DeviceScratch& sc = g_scratch[dev];
	if (bytes > sc.cap) {
		cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
		if (g_default_stream != nullptr && cudaStreamIsCapturing(g_default_stream, &status) == cudaSuccess &&
			status != cudaStreamCaptureStatusNone)
			return nullptr;                      // cannot allocate while capturing: the caller reports PZ_ERR_MEMORY
		const size_t want = bytes < ((size_t)64 << 20) ? ((size_t)64 << 20) : bytes + (bytes >> 2);
		void* fresh = nullptr;
		if (cudaMalloc(&fresh, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
		if (sc.buf) sc.retired.push_back(sc.buf);      // kernels / graphs already enqueued may still use it
		sc.buf = fresh;
		sc.cap = want;
	}
	return sc.buf;
}

// ---------------------------------------------------------------------------------------- launch profiling
namespace {
struct ProfRecord { cudaEvent_t start, stop; int family; double flops, bytes; };
std::vector<ProfRecord> g_prof;
std::vector<cudaEvent_t> g_prof_free;
bool g_prof_enabled = false;

cudaEvent_t prof_event()
{
	if (!g_prof_free.empty()) { cudaEvent_t e = g_prof_free.back(); g_prof_free.pop_back(); return e; }
	cudaEvent_t e;
	cudaEventCreate(&e);
	return e;
}
}  // namespace

bool pz_prof_on() { return g_prof_enabled; }

void pz_prof_begin(int family, cudaStream_t stream, double flops, double bytes)
{
	ProfRecord r{prof_event(), prof_event(), family, flops, bytes};
	cudaEventRecord(r.start, stream);
	g_prof.push_back(r);
}

void pz_prof_end(cudaStream_t stream)
{
	if (!g_prof.empty()) cudaEventRecord(g_prof.back().stop, stream);
}

template <typename T>
__global__ void pz_fill_kernel(T* __restrict__ p, T v, size_t n)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	size_t step = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += step) p[i] = v;
}

template <typename T>
static int pz_fill_any(T* ptr, T value, size_t count, void* stream)
{
	if (count == 0) return PZ_OK;
	// vector path: fill 16-byte words when aligned
	constexpr int per = 16 / sizeof(T);
	size_t head = 0;
	uintptr_t addr = (uintptr_t)ptr;
	if (addr % 16) head = ((16 - addr % 16) / sizeof(T));
	if (head > count) head = count;
	size_t body = (count - head) / per, tail = count - head - body * per;
	int blocks;
	if (head) { pz_fill_kernel<T><<<1, 32, 0, pz_stream(stream)>>>(ptr, value, head); pz_count_launch(1); }
	if (body) {
		uint4 v4;
		T tmp[per];
		for (int i = 0; i < per; i++) tmp[i] = value;
		memcpy(&v4, tmp, 16);
		blocks = (int)(pz_cdiv((int64_t)body, 256) < (int64_t)pz_num_sms() * 16 ? pz_cdiv((int64_t)body, 256) : (int64_t)pz_num_sms() * 16);
		pz_fill_kernel<uint4><<<blocks, 256, 0, pz_stream(stream)>>>((uint4*)(ptr + head), v4, body);
		pz_count_launch(1);
	}
	if (tail) { pz_fill_kernel<T><<<1, 32, 0, pz_stream(stream)>>>(ptr + head + body * per, value, tail); pz_count_launch(1); }
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}


extern "C" {

const char* pz_last_error(void) { return g_err; }
int pz_version(void) { return 100; }
uint64_t pz_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int pz_profile_enable(int on)
{
	g_prof_enabled = on != 0;
	return PZ_OK;
}

// Sums the recorded launches of one family since the last call and clears them (synchronises the device).
int pz_profile_collect(int family, double* total_ms, double* flops, double* bytes, uint64_t* launches)
{
	PZ_CHECK_CUDA(cudaDeviceSynchronize());
	*total_ms = 0.0; *flops = 0.0; *bytes = 0.0; *launches = 0;
	std::vector<ProfRecord> keep;
	for (const ProfRecord& r : g_prof) {
		if (r.family != family && family >= 0) { keep.push_back(r); continue; }
		float ms = 0.0f;
		if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
			*total_ms += ms; *flops += r.flops; *bytes += r.bytes; *launches += 1;
		}
		g_prof_free.push_back(r.start);
		g_prof_free.push_back(r.stop);
	}
	g_prof.swap(keep);
	cudaGetLastError();
	return PZ_OK;
}

int pz_device_count(int* count) { PZ_CHECK_CUDA(cudaGetDeviceCount(count)); return PZ_OK; }
int pz_device_set(int index) { PZ_CHECK_CUDA(cudaSetDevice(index)); return PZ_OK; }
int pz_device_get(int* index) { PZ_CHECK_CUDA(cudaGetDevice(index)); return PZ_OK; }

int pz_device_name(int index, char* buf, int buflen)
{
	cudaDeviceProp prop;
	PZ_CHECK_CUDA(cudaGetDeviceProperties(&prop, index));
	snprintf(buf, (size_t)buflen, "%s", prop.name);
	return PZ_OK;
}

int pz_device_sm_count(int* count) { *count = pz_num_sms(); return PZ_OK; }

int pz_device_cc(int* major, int* minor)
{
	int dev = 0;
	PZ_CHECK_CUDA(cudaGetDevice(&dev));
	PZ_CHECK_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
	PZ_CHECK_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
	return PZ_OK;
}

int pz_device_synchronize(void) { PZ_CHECK_CUDA(cudaDeviceSynchronize()); return PZ_OK; }
int pz_mem_info(size_t* f, size_t* t) { PZ_CHECK_CUDA(cudaMemGetInfo(f, t)); return PZ_OK; }

int pz_malloc(void** ptr, size_t nbytes)
{
	cudaError_t e = cudaMalloc(ptr, nbytes ? nbytes : 1);
	if (e == cudaErrorMemoryAllocation) {
		cudaGetLastError();
		pz_set_error(PZ_ERR_MEMORY, "out of device memory allocating %zu bytes", nbytes);
		return PZ_ERR_MEMORY;
	}
	PZ_CHECK_CUDA(e);
	return PZ_OK;
}

int pz_free(void* ptr) { PZ_CHECK_CUDA(cudaFree(ptr)); return PZ_OK; }
int pz_host_alloc(void** ptr, size_t nbytes) { PZ_CHECK_CUDA(cudaMallocHost(ptr, nbytes ? nbytes : 1)); return PZ_OK; }
int pz_host_free(void* ptr) { PZ_CHECK_CUDA(cudaFreeHost(ptr)); return PZ_OK; }

// ---------------------------------------------------------------------------------------- memory pool
// Size classes follow the reference's float-like binning with two mantissa bits (Allocator.c:29-67):
// a request is rounded up to m * 2^e with m in {4,5,6,7}; the smallest block here is 256 B so that every
// block keeps the 256-byte alignment cudaMalloc gives (128-bit vector access, TMA 16-byte rule).
size_t pz_pool_alloc_size(size_t nbytes)
{
	const size_t minblock = 256;
	if (nbytes <= minblock) return minblock;
	int msb = 63 - __builtin_clzll((unsigned long long)nbytes);
	int shift = msb - 2;
	size_t mant = nbytes >> shift;                 // in [4, 8)
	if (nbytes & (((size_t)1 << shift) - 1)) mant += 1;
	return mant << shift;                           // mant == 8 rolls over to the next exponent, still exact
}

struct PzPool {
	std::mutex mu;
	std::unordered_map<size_t, std::vector<void*>> bins;
	size_t held_blocks = 0, held_bytes = 0, active_blocks = 0, active_bytes = 0;
};

int pz_pool_create(void** pool) { *pool = new PzPool(); return PZ_OK; }

int pz_pool_free_held(void* pool)
{
	PzPool* p = (PzPool*)pool;
	std::lock_guard<std::mutex> lock(p->mu);
	for (auto& kv : p->bins) {
		for (void* ptr : kv.second) cudaFree(ptr);
		kv.second.clear();
	}
	p->held_blocks = 0;
	p->held_bytes = 0;
	return PZ_OK;
}

int pz_pool_destroy(void* pool)
{
	if (!pool) return PZ_OK;
	pz_pool_free_held(pool);
	delete (PzPool*)pool;
	return PZ_OK;
}

int pz_pool_alloc(void* pool, size_t nbytes, void** ptr, size_t* granted)
{
	PzPool* p = (PzPool*)pool;
	size_t sz = pz_pool_alloc_size(nbytes);
	*granted = sz;
	{
		std::lock_guard<std::mutex> lock(p->mu);
		auto it = p->bins.find(sz);
		if (it != p->bins.end() && !it->second.empty()) {
			*ptr = it->second.back();
			it->second.pop_back();
			p->held_blocks -= 1;
			p->held_bytes -= sz;
			p->active_blocks += 1;
			p->active_bytes += sz;
			return PZ_OK;
		}
	}
	int st = pz_malloc(ptr, sz);
	if (st != PZ_OK) return st;      // like the reference, no retry after freeHeld
	std::lock_guard<std::mutex> lock(p->mu);
	p->active_blocks += 1;
	p->active_bytes += sz;
	return PZ_OK;
}

int pz_pool_release(void* pool, void* ptr, size_t granted)
{
	PzPool* p = (PzPool*)pool;
	std::lock_guard<std::mutex> lock(p->mu);
	p->bins[granted].push_back(ptr);
	p->held_blocks += 1;
	p->held_bytes += granted;
	p->active_blocks -= 1;
	p->active_bytes -= granted;
	return PZ_OK;
}

int pz_pool_stats(void* pool, size_t* hb, size_t* hbytes, size_t* ab, size_t* abytes)
{
	PzPool* p = (PzPool*)pool;
	std::lock_guard<std::mutex> lock(p->mu);
	*hb = p->held_blocks; *hbytes = p->held_bytes; *ab = p->active_blocks; *abytes = p->active_bytes;
	return PZ_OK;
}

// ---------------------------------------------------------------------------------------- copies
int pz_memcpy_h2d(void* dst, const void* src, size_t n, void* stream, int async)
{
	if (n == 0) return PZ_OK;
	PZ_CHECK_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, pz_stream(stream)));
	if (!async) PZ_CHECK_CUDA(cudaStreamSynchronize(pz_stream(stream)));
	return PZ_OK;
}

int pz_memcpy_d2h(void* dst, const void* src, size_t n, void* stream, int async)
{
	if (n == 0) return PZ_OK;
	PZ_CHECK_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, pz_stream(stream)));
	if (!async) PZ_CHECK_CUDA(cudaStreamSynchronize(pz_stream(stream)));
	return PZ_OK;
}

int pz_memcpy_d2d(void* dst, const void* src, size_t n, void* stream)
{
	if (n == 0) return PZ_OK;
	PZ_CHECK_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, pz_stream(stream)));
	return PZ_OK;
}

int pz_memcpy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, int kind,
				void* stream)
{
	if (width == 0 || height == 0) return PZ_OK;
	cudaMemcpyKind k = kind == 0 ? cudaMemcpyDeviceToDevice : (kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost);
	PZ_CHECK_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, k, pz_stream(stream)));
	if (kind != 0) PZ_CHECK_CUDA(cudaStreamSynchronize(pz_stream(stream)));
	return PZ_OK;
}

int pz_memset8(void* ptr, uint8_t value, size_t count, void* stream)
{
	if (count == 0) return PZ_OK;
	PZ_CHECK_CUDA(cudaMemsetAsync(ptr, value, count, pz_stream(stream)));
	return PZ_OK;
}

int pz_memset16(void* ptr, uint16_t value, size_t count, void* stream)
{
	if (((value >> 8) & 0xff) == (value & 0xff)) return pz_memset8(ptr, (uint8_t)(value & 0xff), count * 2, stream);
	return pz_fill_any<uint16_t>((uint16_t*)ptr, value, count, stream);
}

int pz_memset32(void* ptr, uint32_t value, size_t count, void* stream)
{
	uint8_t b = value & 0xff;
	if (value == (uint32_t)b * 0x01010101u) return pz_memset8(ptr, b, count * 4, stream);
	return pz_fill_any<uint32_t>((uint32_t*)ptr, value, count, stream);
}

int pz_fill64(void* ptr, uint64_t value, int64_t count, void* stream)
{
	return pz_fill_any<uint64_t>((uint64_t*)ptr, value, (size_t)count, stream);
}

// ---------------------------------------------------------------------------------------- streams / events / graphs
int pz_set_default_stream(void* stream)
{
	g_default_stream = (cudaStream_t)stream;
	return PZ_OK;
}

// Whole-step CUDA graphs: every entry point enqueues on pz_stream(), so a step driven by the unchanged Python operator API can
// be captured once and replayed with a single launch (no per-op host work, no inter-kernel launch gaps).  The caller must have
// run the step at least once before (memory-pool blocks, scratch buffers and kernel attributes exist) and keep the step free
// of host synchronisation.  Scalars (learning rate, batch-norm factor) are frozen at their capture-time values.
int pz_graph_begin(void* stream)
{
	PZ_REQUIRE(stream != nullptr, "graph capture needs a real stream (the legacy default stream cannot be captured)");
	g_default_stream = (cudaStream_t)stream;
	PZ_CHECK_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeRelaxed));
	pz_norm_graph_begin((cudaStream_t)stream);
	return PZ_OK;
}

int pz_graph_end(void* stream, void** exec)
{
	cudaGraph_t graph = nullptr;
	cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &graph);
	if (e != cudaSuccess || graph == nullptr) {
		pz_set_error(PZ_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
		cudaGetLastError();
		return PZ_ERR_CUDA;
	}
	cudaGraphExec_t ge = nullptr;
	e = cudaGraphInstantiate(&ge, graph, 0);
	cudaGraphDestroy(graph);
	if (e != cudaSuccess) {
		pz_set_error(PZ_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(e));
		return PZ_ERR_CUDA;
	}
	*exec = (void*)ge;
	return PZ_OK;
}

int pz_graph_launch(void* exec, void* stream)
{
	PZ_CHECK_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, pz_stream(stream)));
	return PZ_OK;
}

int pz_graph_destroy(void* exec)
{
	PZ_CHECK_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)exec));
	return PZ_OK;
}

int pz_stream_create(void** stream)
{
	cudaStream_t s;
	PZ_CHECK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	*stream = (void*)s;
	return PZ_OK;
}
int pz_stream_destroy(void* stream) { PZ_CHECK_CUDA(cudaStreamDestroy(pz_stream(stream))); return PZ_OK; }
int pz_stream_synchronize(void* stream) { PZ_CHECK_CUDA(cudaStreamSynchronize(pz_stream(stream))); return PZ_OK; }

int pz_event_create(void** event)
{
	cudaEvent_t e;
	PZ_CHECK_CUDA(cudaEventCreate(&e));
	*event = (void*)e;
	return PZ_OK;
}
// an event used only to order streams (fork / join, also under graph capture): no timing
int pz_event_create_sync(void** event)
{
	cudaEvent_t e;
	PZ_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	*event = (void*)e;
	return PZ_OK;
}
int pz_event_destroy(void* event) { PZ_CHECK_CUDA(cudaEventDestroy((cudaEvent_t)event)); return PZ_OK; }
int pz_event_record(void* event, void* stream) { PZ_CHECK_CUDA(cudaEventRecord((cudaEvent_t)event, pz_stream(stream))); return PZ_OK; }
int pz_event_synchronize(void* event) { PZ_CHECK_CUDA(cudaEventSynchronize((cudaEvent_t)event)); return PZ_OK; }
int pz_event_elapsed_ms(void* start, void* stop, float* ms)
{
	PZ_CHECK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
	return PZ_OK;
}
int pz_stream_wait_event(void* stream, void* event)
{
	PZ_CHECK_CUDA(cudaStreamWaitEvent(pz_stream(stream), (cudaEvent_t)event, 0));
	return PZ_OK;
}

}  // extern "C"
