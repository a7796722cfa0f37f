// output_data of the reference (lib/output.py:20-41: generate_output_masks + three skimage.io.imsave calls per page) as
// one asynchronous call: pcs_output_pages enqueues the mask kernel and the PNG encoder for n pages on the context's
// stream and returns; worker threads of the library wait for the encoder, copy ONLY the bytes of the files to
// page-locked memory and write them.  pcs_output_flush waits until every file is on disk.
//
// Why native: the same stages driven from Python (one pinned allocation, three copies, three file writes and a handful
// of GIL hand-overs per page) cost 0.5 ms per page on the caller's thread and 0.6 ms on a single writer thread
// (tools/profile_api.py); the whole drop-in flow needs less than 0.5 ms per page to reach 2 000 pages/s.
#include "common.cuh"

#include <algorithm>
#include <cerrno>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace pcs {

struct OutSlot {
    uint8_t* d_masks = nullptr;            // [3][n][H][Wb]: packed palette indices of the three masks
    size_t masks_bytes = 0;
    uint8_t* d_files = nullptr;            // [3 n][stride]
    size_t files_bytes = 0;
    unsigned long long* d_sizes = nullptr; // [3 n]
    size_t sizes_count = 0;
    cudaEvent_t encoded = nullptr;
    bool busy = false;
};

struct OutJob {
    int slot = 0;
    int files = 0;
    size_t stride = 0;
    std::vector<std::string> paths;        // [files], encoder order (kind-major)
};

struct OutputWriter {
    static constexpr int kMaxSlots = 32;
    int n_slots = 8, n_workers = 4;        // PCSEG_OUTPUT_SLOTS / PCSEG_OUTPUT_WORKERS
    int device = 0;
    OutSlot slots[kMaxSlots];
    int next_slot = 0;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::deque<OutJob> jobs;
    int in_flight = 0;
    bool stop = false;
    std::string err;
    std::vector<std::thread> workers;

    void fail(const std::string& what) {
        std::lock_guard<std::mutex> g(m);
        if (err.empty()) err = what;
    }

    void run() {
        cudaSetDevice(device);
        cudaStream_t st = nullptr;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        uint8_t* h_files = nullptr;
        size_t h_bytes = 0;
        unsigned long long* h_sizes = nullptr;
        size_t h_count = 0;
        for (;;) {
            OutJob job;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return stop || !jobs.empty(); });
                if (jobs.empty()) break;
                job = std::move(jobs.front());
                jobs.pop_front();
            }
            OutSlot& s = slots[job.slot];
            bool ok = true;
            auto check = [&](cudaError_t e, const char* what) {
                if (e != cudaSuccess && ok) { ok = false; fail(std::string(what) + ": " + cudaGetErrorString(e)); }
            };
            check(cudaEventSynchronize(s.encoded), "output: waiting for the encoder");
            if (ok && h_count < (size_t)job.files) {
                if (h_sizes) cudaFreeHost(h_sizes);
                h_count = (size_t)job.files * 2;
                check(cudaHostAlloc(&h_sizes, h_count * sizeof(unsigned long long), cudaHostAllocDefault), "output: cudaHostAlloc");
            }
            if (ok) {
                check(cudaMemcpyAsync(h_sizes, s.d_sizes, (size_t)job.files * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st), "output: sizes");
                check(cudaStreamSynchronize(st), "output: sizes");
            }
            std::vector<size_t> off(job.files + 1, 0);
            if (ok) {
                for (int f = 0; f < job.files; ++f) {
                    if (h_sizes[f] > job.stride) { ok = false; fail("output: encoded file larger than its slot"); break; }
                    off[f + 1] = off[f] + ((size_t)h_sizes[f] + 255) / 256 * 256;
                }
            }
            if (ok && h_bytes < off[job.files]) {
                if (h_files) cudaFreeHost(h_files);
                h_bytes = off[job.files] * 5 / 4 + (1 << 20);
                check(cudaHostAlloc(&h_files, h_bytes, cudaHostAllocDefault), "output: cudaHostAlloc");
                if (!ok) { h_files = nullptr; h_bytes = 0; }
            }
            if (ok) {
                for (int f = 0; f < job.files; ++f)
                    check(cudaMemcpyAsync(h_files + off[f], s.d_files + (size_t)f * job.stride, (size_t)h_sizes[f], cudaMemcpyDeviceToHost, st),
                          "output: file bytes");
                check(cudaStreamSynchronize(st), "output: file bytes");
            }
            {   // the device buffers of the slot are free again: the launching thread may reuse them while we write
                std::lock_guard<std::mutex> g(m);
                s.busy = false;
            }
            cv_done.notify_all();
            if (ok) {
                for (int f = 0; f < job.files; ++f) {
                    FILE* fp = fopen(job.paths[f].c_str(), "wb");
                    const size_t want = (size_t)h_sizes[f];
                    if (!fp || fwrite(h_files + off[f], 1, want, fp) != want) {
                        fail("output: cannot write " + job.paths[f] + ": " + strerror(errno));
                        if (fp) fclose(fp);
                        break;
                    }
                    if (fclose(fp) != 0) { fail("output: cannot write " + job.paths[f] + ": " + strerror(errno)); break; }
                }
            }
            {
                std::lock_guard<std::mutex> g(m);
                --in_flight;
            }
            cv_done.notify_all();
        }
        if (h_files) cudaFreeHost(h_files);
        if (h_sizes) cudaFreeHost(h_sizes);
        if (st) cudaStreamDestroy(st);
    }
};

static OutputWriter* writer_of(pcs_ctx* ctx) {
    if (!ctx->writer) {
        auto* w = new OutputWriter();
        w->device = ctx->device;
        if (const char* e = getenv("PCSEG_OUTPUT_SLOTS")) w->n_slots = std::min(std::max(atoi(e), 1), (int)OutputWriter::kMaxSlots);
        if (const char* e = getenv("PCSEG_OUTPUT_WORKERS")) w->n_workers = std::min(std::max(atoi(e), 1), 32);
        for (int i = 0; i < w->n_workers; ++i) w->workers.emplace_back([w] { w->run(); });
        ctx->writer = w;
    }
    return reinterpret_cast<OutputWriter*>(ctx->writer);
}

void output_writer_destroy(pcs_ctx* ctx) {
    auto* w = reinterpret_cast<OutputWriter*>(ctx->writer);
    if (!w) return;
    {
        std::lock_guard<std::mutex> g(w->m);
        w->stop = true;
    }
    w->cv_job.notify_all();
    for (auto& t : w->workers) t.join();
    for (auto& s : w->slots) {
        if (s.d_masks) cudaFree(s.d_masks);
        if (s.d_files) cudaFree(s.d_files);
        if (s.d_sizes) cudaFree(s.d_sizes);
        if (s.encoded) cudaEventDestroy(s.encoded);
    }
    delete w;
    ctx->writer = nullptr;
}

int output_flush(pcs_ctx* ctx) {
    auto* w = reinterpret_cast<OutputWriter*>(ctx->writer);
    if (!w) return PCS_OK;
    std::unique_lock<std::mutex> lk(w->m);
    w->cv_done.wait(lk, [&] { return w->in_flight == 0; });
    if (!w->err.empty()) {
        std::string e;
        e.swap(w->err);
        lk.unlock();
        return set_err(ctx, PCS_ERR_IO, "%s", e.c_str());
    }
    return PCS_OK;
}

int output_pages(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary, int n, int H, int W, const uint8_t* lut, int n_lut,
                 const char* const* paths) {
    if (n <= 0 || n > 64) return set_err(ctx, PCS_ERR_ARG, "output_pages: 1..64 pages per call, got %d", n);
    if (n_lut > 255) return set_err(ctx, PCS_ERR_ARG, "output_pages: at most 255 LUT rows (one palette entry is black)");
    // the masks hold n_lut + 1 distinct colours: indexed-colour PNGs (png.cu), 2 bits per pixel for the default colour map
    const int ncolors = n_lut + 1;
    uint8_t palette[768] = {};
    memcpy(palette, lut, (size_t)n_lut * 3);                  // entry n_lut stays black
    int level = 1;
    size_t bound = png_indexed_file_bytes(H, W, ncolors, 1);
    if (!bound) { level = 0; bound = png_indexed_file_bytes(H, W, ncolors, 0); }
    if (!bound) return set_err(ctx, PCS_ERR_ARG, "output_pages: masks of %d x %d cannot be written as PNG", H, W);
    const size_t stride = (bound + 255) / 256 * 256;
    OutputWriter* w = writer_of(ctx);
    int si;
    {
        std::unique_lock<std::mutex> lk(w->m);
        if (!w->err.empty()) {
            std::string e;
            e.swap(w->err);
            lk.unlock();
            return set_err(ctx, PCS_ERR_IO, "%s", e.c_str());
        }
        si = w->next_slot;
        w->next_slot = (si + 1) % w->n_slots;
        w->cv_done.wait(lk, [&] { return !w->slots[si].busy; });
        w->slots[si].busy = true;
        ++w->in_flight;
    }
    OutSlot& s = w->slots[si];
    auto release = [&](int rc) {
        {
            std::lock_guard<std::mutex> g(w->m);
            s.busy = false;
            --w->in_flight;
        }
        w->cv_done.notify_all();
        return rc;
    };
    const int files = 3 * n;
    const int depth = png_index_depth(ncolors), Wb = (W * depth + 7) / 8;
    const size_t plane = (size_t)n * H * Wb;                  // packed palette indices of one kind of mask
    auto grow = [&](void** p, size_t* have, size_t need) -> int {
        if (*have >= need) return PCS_OK;
        if (*p) cudaFree(*p);          // the slot is idle: nothing in flight touches its buffers
        *p = nullptr;
        *have = 0;
        const size_t want = need + need / 4;
        if (cudaMalloc(p, want) != cudaSuccess) {
            cudaGetLastError();
            return set_err(ctx, PCS_ERR_NOMEM, "output_pages: cudaMalloc of %zu bytes failed", want);
        }
        *have = want;
        return PCS_OK;
    };
    int rc;
    if ((rc = grow(reinterpret_cast<void**>(&s.d_masks), &s.masks_bytes, 3 * plane + 16)) != PCS_OK) return release(rc);
    if ((rc = grow(reinterpret_cast<void**>(&s.d_files), &s.files_bytes, (size_t)files * stride)) != PCS_OK) return release(rc);
    size_t sizes_bytes = s.sizes_count * sizeof(unsigned long long);
    if ((rc = grow(reinterpret_cast<void**>(&s.d_sizes), &sizes_bytes, (size_t)files * sizeof(unsigned long long))) != PCS_OK) return release(rc);
    s.sizes_count = sizes_bytes / sizeof(unsigned long long);
    if (!s.encoded && cudaEventCreateWithFlags(&s.encoded, cudaEventDisableTiming) != cudaSuccess)
        return release(set_err(ctx, PCS_ERR_CUDA, "output_pages: cudaEventCreate failed"));
    if ((rc = launch_mask_indices(ctx, d_labels, d_binary, n, H, W, n_lut, s.d_masks)) != PCS_OK) return release(rc);
    if ((rc = launch_png_encode_indexed(ctx, s.d_masks, files, H, W, palette, ncolors, level, s.d_files, stride, s.d_sizes)) != PCS_OK)
        return release(rc);
    if (cudaEventRecord(s.encoded, ctx->stream) != cudaSuccess) return release(set_err(ctx, PCS_ERR_CUDA, "output_pages: cudaEventRecord failed"));
    OutJob job;
    job.slot = si;
    job.files = files;
    job.stride = stride;
    job.paths.resize(files);
    for (int p = 0; p < n; ++p)
        for (int k = 0; k < 3; ++k) job.paths[(size_t)k * n + p] = paths[3 * p + k];      // encoder order: kind-major
    {
        std::lock_guard<std::mutex> g(w->m);
        w->jobs.push_back(std::move(job));
    }
    w->cv_job.notify_one();
    return PCS_OK;
}

}  // namespace pcs
