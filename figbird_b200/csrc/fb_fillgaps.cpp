// fb_fillgaps.cpp -- fb_fillgaps_main(): the drop-in for the `FillGaps` executable.
//
// Same 15 positional arguments, same input files, same output files as FillGaps.cpp:371-947 + the worker
// processes it spawns (Figbird.cpp main :6909-7508).  What is different by design:
//   * no worker processes, no run-time compilation: the model is learned once, gaps are sharded
//     cost-balanced over the visible GPUs (FIGBIRD_GPUS), one engine context per GPU, no collective;
//   * every gap runs its sequential control logic on a fiber of a small worker pool; the device requests of all
//     gaps in flight on a lane are merged into one fb_em_run per tick (LaneQueue below);
//   * draw.txt is written once, in the order the reference's per-worker files would have been concatenated (referenceDrawOrder below).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <time.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>

#include "fb_gapfiller.h"
#include "fb_io.h"

namespace fb {

// ---------------------------------------------------------------------------------------------------------
// Lane scheduler.  A lane = one engine context + a few worker threads.  Every gap runs the reference's
// *sequential* control (GapFill::run) on its own fiber (ucontext, 256 KB stack): a device request parks the
// fiber instead of blocking an OS thread, so thousands of gaps can be in flight on a handful of threads.
// A worker owns its fibers for their whole life (no migration).  When all of a worker's fibers are parked it
// hands their requests to the lane; the last worker to arrive flushes them as ONE fb_em_run, after which every
// worker resumes its fibers and each fiber reads its own results in place in the pinned arena (which stays valid
// until the next flush, and that needs every worker -- hence this fiber -- to arrive again).
// ---------------------------------------------------------------------------------------------------------
struct Fiber {
    ucontext_t ctx, *ret = nullptr;
    void* stack = nullptr; size_t stackBytes = 0;
    GapFill* fill = nullptr; int bidx = 0, gap = 0;
    DeviceQueue* dev = nullptr;
    GapResult result; std::string error;
    const std::vector<ItemSpec>* reqItems = nullptr;
    std::vector<const FbItemOut*> outs;
    bool done = false, failed = false;
};
static thread_local Fiber* tlsFiber = nullptr;

class LaneQueue : public DeviceQueue {
public:
    LaneQueue(fb_ctx* ctx, int workers, std::function<std::string()> beforeFirstRun) : ctx_(ctx), running_(workers), beforeFirstRun_(std::move(beforeFirstRun)) {}
    // called on a fiber: park until the lane has run the request
    void submit(int, const std::vector<ItemSpec>& items, std::vector<ItemResult>& results) override {
        Fiber* f = tlsFiber;
        f->reqItems = &items; f->outs.clear();
        swapcontext(&f->ctx, f->ret);
        if (f->failed) throw std::runtime_error(err_);
        results.clear(); results.reserve(f->outs.size());
        for (const FbItemOut* h : f->outs) {
            results.emplace_back();
            ItemResult& res = results.back(); const unsigned char* b = (const unsigned char*)h;
            res.calls = h->calls; res.compCount = h->comp_count; res.flags = h->flags; res.nReads = h->n_reads;
            res.candLen = h->cand_len; res.nSlots = h->n_slots; res.placements = h->placements;
            res.p1max = (const double*)(b + h->off_p1max); res.p2max = (const double*)(b + h->off_p2max); res.pos2 = (const int32_t*)(b + h->off_pos2);
            res.soft = b + h->off_soft; res.hard = b + h->off_hard; res.cov = (const int32_t*)(b + h->off_cov);
            res.counts = h->off_counts >= 0 ? (const double*)(b + h->off_counts) : nullptr;
        }
        f->reqItems = nullptr;
    }
    // called by a worker thread whose fibers are all parked: returns when their requests have been run
    void exchange(const std::vector<Fiber*>& parked) {
        std::unique_lock<std::mutex> lk(mu_);
        for (Fiber* f : parked) reqs_.push_back(f);
        arrived_++;
        if (arrived_ >= running_) flush(lk);
        else { const int64_t gen = gen_; cv_.wait(lk, [&] { return gen_ != gen; }); }
    }
    void workerExit() {
        std::unique_lock<std::mutex> lk(mu_);
        running_--;
        if (running_ > 0 && arrived_ >= running_) flush(lk);
    }
    int64_t ticks() const { return ticks_; }
    double engineSeconds() const { return tEngine_; }
    int activeGaps() const override { return lastGaps_.load(std::memory_order_relaxed); }
    void dumpTickLog(int lane) const {      // FIGBIRD_TICK_LOG: start (s since the lane's first tick), engine call (ms), items, gaps of every tick
        if (!tickLog_) return;
        std::string o = "ticklog lane " + std::to_string(lane) + " ticks " + std::to_string(log_.size()) + ":";
        for (size_t i = 0; i < log_.size(); i++) { if (log_.size() > 60 && i >= 25 && i + 35 < log_.size()) { if (i == 25) o += " ..."; continue; } char b[96]; snprintf(b, sizeof b, " [%.3f %.1fms %d/%d]", log_[i].t, 1e3 * log_[i].d, log_[i].items, log_[i].gaps); o += b; }
        fprintf(stderr, "%s\n", o.c_str());
    }

private:
    void flush(std::unique_lock<std::mutex>&) {
        std::vector<Fiber*> batch; batch.swap(reqs_);
        arrived_ = 0;
        lastGaps_.store((int)batch.size(), std::memory_order_relaxed);
        std::vector<FbWorkItem> wi;
        for (Fiber* f : batch) for (const ItemSpec& s : *f->reqItems) {
            FbWorkItem w{}; w.kind = s.kind; w.gap = f->bidx; w.cand_len = s.candLen; w.max_rounds = s.maxRounds; w.flags = s.flags;
            w.comp_count_in = s.compIn; w.counts_in = s.countsIn.empty() ? nullptr : s.countsIn.data();
            w.string_in = s.stringIn.empty() ? nullptr : s.stringIn.data();
            wi.push_back(w);
        }
        outs_.assign(wi.size(), nullptr);
        if (beforeFirstRun_) {      // the model: awaited and uploaded here, in front of the lane's first engine call
            const std::string e = beforeFirstRun_();
            beforeFirstRun_ = nullptr;
            if (!e.empty()) { failed_ = true; err_ = e; }
        }
        auto c0 = std::chrono::steady_clock::now();
        fb_status st = (wi.empty() || failed_) ? FB_OK : fb_em_run(ctx_, wi.data(), (int32_t)wi.size(), outs_.data());
        auto c1 = std::chrono::steady_clock::now();
        tEngine_ += std::chrono::duration<double>(c1 - c0).count();
        if (tickLog_) { if (ticks_ == 0) t0_ = c0; log_.push_back({std::chrono::duration<double>(c0 - t0_).count(), std::chrono::duration<double>(c1 - c0).count(), (int)wi.size(), (int)batch.size()}); }
        ticks_++;
        if (st != FB_OK && !failed_) { failed_ = true; err_ = std::string("fb_em_run failed: ") + fb_last_error(ctx_); }
        size_t k = 0;
        for (Fiber* f : batch) {
            const size_t n = f->reqItems->size();
            f->failed = failed_;
            if (!failed_) f->outs.assign(outs_.begin() + k, outs_.begin() + k + n);
            k += n;
        }
        gen_++;
        cv_.notify_all();
    }
    fb_ctx* ctx_;
    std::mutex mu_; std::condition_variable cv_;
    std::vector<Fiber*> reqs_;
    std::vector<const FbItemOut*> outs_;
    int running_, arrived_ = 0;
    std::function<std::string()> beforeFirstRun_;
    int64_t gen_ = 0;
    bool failed_ = false; std::string err_;
    int64_t ticks_ = 0;
    double tEngine_ = 0;
    std::atomic<int> lastGaps_{1 << 30};
    struct TickRec { double t, d; int items, gaps; };
    const bool tickLog_ = getenv("FIGBIRD_TICK_LOG") != nullptr;
    std::chrono::steady_clock::time_point t0_;
    std::vector<TickRec> log_;
};

static void fiberEntry(unsigned lo, unsigned hi) {
    Fiber* f = (Fiber*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    try { f->result = f->fill->run(*f->dev, f->bidx); }
    catch (const std::exception& e) { f->error = e.what(); if (f->error.empty()) f->error = "gap worker failed"; }
    catch (...) { f->error = "gap worker failed"; }
    f->done = true;
    // returning switches to uc_link (the worker's context)
}

// One worker thread of a lane: keeps up to `cap` gaps in flight on fibers, taking gaps from the lane's list.
struct LaneWork {
    const std::vector<int>* mine = nullptr;
    std::vector<std::unique_ptr<GapFill>>* fills = nullptr;
    std::vector<GapResult>* results = nullptr;
    std::atomic<int> next{0};
    std::mutex emu; std::string error;
};
static void laneWorker(LaneQueue& q, LaneWork& lw, int cap) {
    constexpr size_t kStack = 256 * 1024;
    ucontext_t self;
    std::vector<Fiber*> active, parked;
    std::vector<void*> freeStacks;
    auto resume = [&](Fiber* f) { tlsFiber = f; f->ret = &self; swapcontext(&self, &f->ctx); tlsFiber = nullptr; };
    auto retire = [&](Fiber* f) {
        if (!f->error.empty()) { std::lock_guard<std::mutex> l(lw.emu); if (lw.error.empty()) lw.error = f->error; }
        else (*lw.results)[f->gap] = std::move(f->result);
        // the gap's working set (reads, tables, strings) is freed here, beside the kernels of the other lane, instead of in one serial
        // sweep after the last gap (0.1-0.2 s per call at 10 k gaps); nothing reads it once run() has returned
        (*lw.fills)[f->gap].reset();
        freeStacks.push_back(f->stack);
        delete f;
    };
    auto topUp = [&]() {
        while ((int)active.size() < cap) {
            const int i = lw.next++;
            if (i >= (int)lw.mine->size()) break;
            Fiber* f = new Fiber();
            f->gap = (*lw.mine)[i]; f->bidx = i; f->fill = (*lw.fills)[f->gap].get(); f->dev = &q;
            if (!freeStacks.empty()) { f->stack = freeStacks.back(); freeStacks.pop_back(); }
            else f->stack = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
            if (f->stack == MAP_FAILED) { f->stack = nullptr; f->error = "cannot allocate a fiber stack"; std::lock_guard<std::mutex> l(lw.emu); if (lw.error.empty()) lw.error = f->error; delete f; break; }
            f->stackBytes = kStack;
            getcontext(&f->ctx);
            f->ctx.uc_stack.ss_sp = f->stack; f->ctx.uc_stack.ss_size = kStack; f->ctx.uc_link = &self;
            const uintptr_t p = (uintptr_t)f;
            makecontext(&f->ctx, (void (*)())fiberEntry, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
            resume(f);
            if (f->done) retire(f); else active.push_back(f);
        }
    };
    topUp();
    while (!active.empty()) {
        q.exchange(active);
        parked.swap(active); active.clear();
        for (Fiber* f : parked) { resume(f); if (f->done) retire(f); else active.push_back(f); }
        parked.clear();
        topUp();
    }
    for (void* s : freeStacks) munmap(s, kStack);
    q.workerExit();
}

static bool parseArgs(int argc, const char* const* argv, Args& a) {
    if (argc < 16) return false;
    a.draft = argv[1]; a.maxDistance = atoi(argv[2]); a.readLength = atoi(argv[3]); a.scriptItr = atoi(argv[4]);
    a.partialFlag = atoi(argv[5]); a.unmapped = atoi(argv[6]); a.numThreads = atoi(argv[7]); a.myout = argv[8];
    a.tmpDir = argv[9]; a.gapsDir = argv[10]; a.negOverlap = atoi(argv[11]); a.partialReadLen = atoi(argv[12]);
    a.trim = atoi(argv[13]); a.setInputMean = atoi(argv[14]); a.insertSizeMean = atoi(argv[15]);
    return true;
}

template <class F>
static void parallelFor(int n, int threads, F f) {
    threads = std::max(1, std::min(threads, n));
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    std::mutex emu; std::string error; bool failed = false;
    for (int t = 0; t < threads; t++) th.emplace_back([&] {
        try { for (int i; (i = next++) < n;) f(i); }
        catch (const std::exception& e) { std::lock_guard<std::mutex> l(emu); if (!failed) { failed = true; error = e.what(); } next = n; }
        catch (...) { std::lock_guard<std::mutex> l(emu); if (!failed) { failed = true; error = "worker failed"; } next = n; }
    });
    for (auto& t : th) t.join();
    if (failed) throw std::runtime_error(error);      // rethrown on the caller's thread (fb_fillgaps_main turns it into exit status 1)
}

// Engine contexts survive fb_fillgaps_main: an in-process caller (one FillGaps call per pipeline iteration) pays stream /
// buffer / pinned-arena creation once per (device, lane).  Leaked on purpose at process exit.
struct CtxPool { std::mutex mu; std::vector<std::pair<long, fb_ctx*>> idle; };      // key = device * 1024 + lane: a lane gets its own context back (arenas already sized for its shard)
static CtxPool& ctxPool() { static CtxPool* p = new CtxPool; return *p; }
static fb_ctx* acquireCtx(int device, int lane, std::string& err) {
    {
        CtxPool& P = ctxPool(); std::lock_guard<std::mutex> l(P.mu);
        const long key = (long)device * 1024 + lane;
        for (size_t i = 0; i < P.idle.size(); i++) if (P.idle[i].first == key) { fb_ctx* c = P.idle[i].second; P.idle.erase(P.idle.begin() + i); return c; }
    }
    fb_ctx* ctx = nullptr;
    if (fb_ctx_create(device, &ctx) != FB_OK || !ctx) {
        err = std::string("fb_ctx_create failed on device ") + std::to_string(device) + ": " + (ctx ? fb_last_error(ctx) : "no context");
        if (ctx) fb_ctx_destroy(ctx);
        return nullptr;
    }
    return ctx;
}
static void releaseCtx(int device, int lane, fb_ctx* ctx) { CtxPool& P = ctxPool(); std::lock_guard<std::mutex> l(P.mu); P.idle.emplace_back((long)device * 1024 + lane, ctx); }

// Lanes: independent (context, batch queue, worker threads, fibers) groups on one GPU.  Two throughput lanes per GPU alternate --
// the host replay of one overlaps the kernels of the other.  A third, *latency* lane per GPU takes the gaps whose control is a long
// chain of single-round requests (unmapped mode, N-run > 400: up to 200 host-driven rounds with border updates, Figbird.cpp:4029-4376):
// its ticks are tiny and run beside the big batches instead of queueing one round behind each of them.
struct LaneSpec { int device; int kind; };      // kind 0: throughput, 1: latency
static std::vector<LaneSpec> visibleLanes() {
    std::vector<int> d;
    const char* e = getenv("FIGBIRD_GPUS");
    if (e && *e) { for (const char* p = e; *p;) { d.push_back(atoi(p)); while (*p && *p != ',') p++; if (*p) p++; } }
    if (d.empty()) d.push_back(0);
    int lanes = 2;
    if (const char* l = getenv("FIGBIRD_LANES")) lanes = std::max(1, atoi(l));
    const bool latency = !(getenv("FIGBIRD_LATENCY_LANE") && atoi(getenv("FIGBIRD_LATENCY_LANE")) == 0);
    std::vector<LaneSpec> out;
    for (int x : d) { for (int i = 0; i < lanes; i++) out.push_back(LaneSpec{x, 0}); if (latency) out.push_back(LaneSpec{x, 1}); }
    return out;
}

// The order in which the reference's draw.txt lists the gaps: FillGaps deals the gaps to num_threads workers (N-runs of at most
// 400 bases round-robin, the longer ones to the workers with the most room left; FillGaps.cpp:456-649), every worker writes its
// gaps in ascending order (writeGapLoad :313-334, Figbird.cpp:7283-7317) and the workers' files are concatenated in worker order
// (mergeFiles :222-258).  The sort of the workers by remaining room is std::sort with the reference's comparator on the
// reference's container, i.e. the same unstable order for equal keys.
// firstWorker = how many entries at the front of the order belong to worker 0, workers = number of worker files that are merged.
static std::vector<int> referenceDrawOrder(const std::string& tmpDir, int totGaps, int numThreads, int& firstWorker, int& workers) {
    std::vector<int> order;
    firstWorker = 0; workers = 1;
    if (totGaps <= 0) return order;
    int T = numThreads;
    std::vector<std::vector<int>> alloc;
    if (totGaps <= T || T < 1) {      // one gap per worker (FillGaps.cpp:460-464, 497-505)
        for (int i = 0; i < totGaps; i++) order.push_back(i);
        firstWorker = 1; workers = T < 1 ? 1 : totGaps;
        return order;
    }
    const float tf = (float)(totGaps * 1.0 / T);
    int per = (int)tf;
    if (tf - float(per) > 0) per++;
    std::vector<int> small, large;
    {
        FILE* f = fopen((tmpDir + "gapInfo.txt").c_str(), "r");
        if (!f) { for (int i = 0; i < totGaps; i++) order.push_back(i); return order; }
        char line[1024]; int cnt = 0;
        while (fgets(line, sizeof line, f)) {
            char* t = strtok(line, "\t"); t = strtok(nullptr, "\t"); t = strtok(nullptr, "\t\n");
            const int gaplen = t ? atoi(t) : 0;
            (gaplen > 400 ? large : small).push_back(cnt);
            cnt++;
        }
        fclose(f);
    }
    alloc.assign((size_t)T, std::vector<int>());
    const int nSmall = (int)small.size(), nLarge = (int)large.size();
    if (nSmall <= T) { for (int i = 0; i < nSmall; i++) alloc[(size_t)i].push_back(small[(size_t)i]); }
    else { for (int k = 0; k < nSmall; k++) alloc[(size_t)(k % T)].push_back(small[(size_t)k]); }
    std::vector<std::vector<int>> rem((size_t)T);
    for (int i = 0; i < T; i++) rem[(size_t)i] = std::vector<int>{i, per - (int)alloc[(size_t)i].size()};
    std::sort(rem.begin(), rem.end(), [](const std::vector<int>& a, const std::vector<int>& b) { return a[1] > b[1]; });
    int k = 0;
    if (nLarge <= T) { for (int i = 0; i < T && k < nLarge; i++) alloc[(size_t)rem[(size_t)i][0]].push_back(large[(size_t)k++]); }
    else {
        for (int i = 0; i < T && k < nLarge; i++)
            for (int j = 0; j < rem[(size_t)i][1] && k < nLarge; j++) alloc[(size_t)rem[(size_t)i][0]].push_back(large[(size_t)k++]);
    }
    for (auto& a : alloc) { std::sort(a.begin(), a.end()); for (int g : a) order.push_back(g); }
    firstWorker = (int)alloc[0].size(); workers = T;
    return order;
}

struct RunStats { double tLoad = 0, tModel = 0, tPrep = 0, tFill = 0, tWrite = 0, tEngine = 0, tCopy = 0, tCtx = 0, tWorkers = 0, cpuWorkers = 0; int64_t refPlacements = 0; FbCounters dev{}; int64_t ticks = 0; };

int fillgapsMain(int argc, const char* const* argv) {
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    Args a;
    if (!parseArgs(argc, argv, a)) { fprintf(stderr, "usage: fillgaps <15 FillGaps arguments> (FillGaps.cpp:419-433)\n"); return 1; }
    RunStats rs;
    auto t0 = clk::now();
    std::vector<GapRecord> gaps; int totGaps = 0;
    if (!loadGapRecords(a.tmpDir, gaps, totGaps)) { fprintf(stderr, "Couldn't open gapinfo in Fillgaps.cpp\n"); return 1; }
    const bool quiet = getenv("FIGBIRD_QUIET") != nullptr;
    if (!quiet) printf("Total # of gaps = %d\n", totGaps);
    Scaffolds sc;
    // the model is learned on its own thread while the scaffolds and the per-gap inputs are read, encoded and uploaded; nothing
    // before the first engine call needs it.  It maps myout.sam and finds the line starts first, then waits for the scaffolds
    // (their lengths enter the statistics).
    Model model; std::string err; bool modelOk = false; double modelSecs = 0;
    std::mutex modelMu; std::condition_variable modelCv; bool modelDone = false;
    int scState = 0;      // 0: loading, 1: loaded, -1: failed (under modelMu)
    auto waitScaffolds = [&]() -> bool { std::unique_lock<std::mutex> l(modelMu); modelCv.wait(l, [&] { return scState != 0; }); return scState > 0; };
    std::thread modelThread([&] {
        auto m0 = clk::now();
        bool ok = learnModel(a, sc, model, err, waitScaffolds);
        std::lock_guard<std::mutex> l(modelMu); modelOk = ok; modelSecs = secs(m0, clk::now()); modelDone = true; modelCv.notify_all();
    });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } modelJoiner{modelThread};
    {
        const bool scOk = loadScaffolds(a.draft, sc);
        { std::lock_guard<std::mutex> l(modelMu); scState = scOk ? 1 : -1; }
        modelCv.notify_all();
        if (!scOk) { printf("Can't open contig file\n"); return 1; }
    }
    auto t1 = clk::now();
    auto waitModel = [&]() -> bool { std::unique_lock<std::mutex> l(modelMu); modelCv.wait(l, [&] { return modelDone; }); return modelOk; };
    auto t2 = t1;
    if (getenv("FIGBIRD_DUMP_MODEL") && waitModel()) {
        FILE* fm = fopen(getenv("FIGBIRD_DUMP_MODEL"), "w");
        if (fm) {
            fprintf(fm, "mean %.17g leftSD %.17g rightSD %.17g tmin %d tmax %d cutoff %d maxins %d maxread %d\n", model.insertSizeMean, model.leftSD, model.rightSD,
                    model.insertThresholdMin, model.insertThresholdMax, model.gapProbCutOff, model.maxInsertSize, model.maxReadLength);
            for (int i = 0; i < model.maxReadLength; i++) fprintf(fm, "pos %d %.17g %.17g %.17g\n", i, model.errorPosDist[i], model.inPosDist[i], model.delPosDist[i]);
            for (int i = 0; i < 5; i++) fprintf(fm, "etp %d %.17g %.17g %.17g %.17g %.17g\n", i, model.errorTypeProbs[i][0], model.errorTypeProbs[i][1], model.errorTypeProbs[i][2], model.errorTypeProbs[i][3], model.errorTypeProbs[i][4]);
            for (int i = 0; i < model.maxInsertSize; i++) if (i < 1200 || i % 97 == 0) fprintf(fm, "pdf %d %.17g\n", i, model.insertPdfSmoothed[i]);
            fclose(fm);
        }
    }

    // ---- per-gap inputs + host-only analysis
    const int nG = (int)gaps.size();
    std::vector<std::unique_ptr<GapFill>> fills(nG);
    const int ioThreads = std::max(1, std::min(a.numThreads > 0 ? a.numThreads : 1, 64));
    int hostThreads = std::max(ioThreads, (int)std::thread::hardware_concurrency());
    if (const char* e = getenv("FIGBIRD_HOST_THREADS")) hostThreads = std::max(1, atoi(e));     // several ranks on one host share its cores
    // per-gap inputs: the text files of the contract, or -- only when the caller asks for it (FIGBIRD_CONTAINER=1) and fb_preprocess_main
    // left them -- the containers that hold the same bytes in one mapped file each
    GapContainer partialBox, unmappedBox;
    if (const char* e = getenv("FIGBIRD_CONTAINER")) if (atoi(e) != 0) {
        partialBox.open(a.gapsDir + "partial_gaps.fbc", 1, (size_t)totGaps);
        if (a.unmapped == 1) unmappedBox.open(a.gapsDir + "gaps.fbc", 2, (size_t)totGaps);
    }
    std::atomic<long long> tIo(0), tCtor(0), tPrep(0);
    parallelFor(nG, hostThreads, [&](int g) {
        auto x0 = clk::now();
        GapInput in; in.rec = gaps[g];
        size_t tn = 0;
        if (partialBox.valid()) { const char* t = partialBox.text((size_t)g, tn); loadPartialText(t, tn, in.partial); in.partialExists = true; }
        else loadPartial(a.gapsDir + "partial_gaps_" + std::to_string(g) + ".sam", in.partial, in.partialExists);
        if (a.unmapped == 1) {
            if (unmappedBox.valid()) { const char* t = unmappedBox.text((size_t)g, tn); loadUnmappedText(t, tn, in.unm, in.unmPairCount); }
            else loadUnmapped(a.gapsDir + "gaps_" + std::to_string(g) + ".sam", a.readLength, in.unm, in.unmPairCount);
        }
        auto x1 = clk::now();
        fills[g].reset(new GapFill(a, model, sc, std::move(in)));
        auto x2 = clk::now();
        fills[g]->prepare();
        auto x3 = clk::now();
        tIo += std::chrono::duration_cast<std::chrono::nanoseconds>(x1 - x0).count(); tCtor += std::chrono::duration_cast<std::chrono::nanoseconds>(x2 - x1).count(); tPrep += std::chrono::duration_cast<std::chrono::nanoseconds>(x3 - x2).count();
    });
    if (getenv("FIGBIRD_PREP_TIMING")) fprintf(stderr, "prepare: threads %d, summed over threads: load %.3f s, ctor %.3f s, prepare %.3f s\n", hostThreads, tIo.load() * 1e-9, tCtor.load() * 1e-9, tPrep.load() * 1e-9);
    auto t3 = clk::now();

    // ---- shard gaps over GPUs: longest-processing-time-first on the cost estimate
    const std::vector<LaneSpec> laneSpecs = visibleLanes();
    std::vector<int> devs; for (auto& l : laneSpecs) devs.push_back(l.device);      // device of every lane
    const int nD = (int)devs.size();
    std::vector<int> gpuList; for (int x : devs) if (std::find(gpuList.begin(), gpuList.end(), x) == gpuList.end()) gpuList.push_back(x);
    const int nGpus = (int)gpuList.size();
    std::vector<std::vector<int>> shard(nD);
    {
        // two levels, both longest-processing-time-first on the cost estimate: gaps -> GPUs, then a GPU's gaps -> its lanes
        std::vector<int> order(nG); for (int i = 0; i < nG; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return fills[x]->prepared().cost > fills[y]->prepared().cost; });
        std::vector<double> gload(nGpus, 0);
        std::vector<std::vector<int>> gshard(nGpus);
        for (int g : order) { int d = (int)(std::min_element(gload.begin(), gload.end()) - gload.begin()); gshard[d].push_back(g); gload[d] += fills[g]->prepared().cost + 1; }
        for (int gi = 0; gi < nGpus; gi++) {
            std::vector<int> thr, lat;
            for (int d = 0; d < nD; d++) if (devs[d] == gpuList[gi]) (laneSpecs[d].kind == 1 ? lat : thr).push_back(d);
            std::vector<double> load(thr.size(), 0);
            for (int g : gshard[gi]) {
                if (!lat.empty() && fills[g]->prepared().sequential) { shard[lat[0]].push_back(g); continue; }
                int k = (int)(std::min_element(load.begin(), load.end()) - load.begin());
                shard[thr[k]].push_back(g); load[k] += fills[g]->prepared().cost + 1;
            }
        }
    }
    std::vector<GapResult> results(nG);
    std::vector<char> onlyGap(nG, 1);
    if (const char* og = getenv("FIGBIRD_ONLY_GAPS")) {   // debugging aid: fill only the listed gaps, leave the rest as N
        std::fill(onlyGap.begin(), onlyGap.end(), 0);
        for (const char* q = og; *q;) { int v = atoi(q); if (v >= 0 && v < nG) onlyGap[v] = 1; while (*q && *q != ',') q++; if (*q) q++; }
        for (int g = 0; g < nG; g++) if (!onlyGap[g]) { results[g].gapStringLength = gaps[g].gapLength; results[g].gapString.assign(gaps[g].gapLength, 'N'); }
        for (auto& sh : shard) { std::vector<int> keep; for (int g : sh) if (onlyGap[g]) keep.push_back(g); sh.swap(keep); }
    }
    std::vector<std::string> devErr(nD);
    std::vector<FbCounters> devCtr(nD); std::vector<int64_t> devTicks(nD, 0); std::vector<double> devEng(nD, 0), devCopy(nD, 0), devCtx(nD, 0), devWork(nD, 0), devCpu(nD, 0);
    std::vector<std::thread> devThreads;
    for (int d = 0; d < nD; d++) devThreads.emplace_back([&, d] {
        const std::vector<int>& mine = shard[d];
        if (mine.empty()) return;
        auto d0 = clk::now();
        int laneOf = 0; for (int e2 = 0; e2 < d; e2++) if (devs[e2] == devs[d]) laneOf++;
        fb_ctx* ctx = acquireCtx(devs[d], laneOf, devErr[d]);
        if (!ctx) return;
        fb_ctx_set_latency_critical(ctx, laneSpecs[d].kind == 1 ? 1 : 0);
        FbCounters ctr0{}; fb_get_counters(ctx, &ctr0);
        // batch of this shard
        std::vector<FbGap> fg; std::vector<int32_t> rlen, rmate, pileL, pileR; std::vector<int64_t> roff; std::vector<uint8_t> rfl, rjlo, rjcut, codes, flank;
        for (int g : mine) {
            const PreparedGap& p = fills[g]->prepared();
            FbGap G{}; G.gap_start = p.gapStart; G.mode = p.mode; G.orig_len = p.origLen; G.n_reads = p.supported ? (int)p.readLen.size() : 0;
            G.read_begin = (int)rlen.size(); G.flank_len = p.flankLen; G.flank_begin = (int)flank.size(); G.pile_len = p.pileLen; G.pile_begin = (int)(pileL.size() / 4);
            if (p.supported) {
                for (size_t q = 0; q < p.readLen.size(); q++) {
                    rlen.push_back(p.readLen[q]); rmate.push_back(p.readMate[q]); rfl.push_back(p.readFlags[q]); rjlo.push_back(p.readJlo[q]); rjcut.push_back(p.readJcut[q]);
                    roff.push_back((int64_t)codes.size()); codes.insert(codes.end(), p.readCodes[q].begin(), p.readCodes[q].end());
                }
                flank.insert(flank.end(), p.flank.begin(), p.flank.end());
                pileL.insert(pileL.end(), p.pileL.begin(), p.pileL.end()); pileR.insert(pileR.end(), p.pileR.begin(), p.pileR.end());
            } else { G.flank_len = 0; G.pile_len = 0; }
            fg.push_back(G);
        }
        // keep arrays non-null for empty shards
        if (rlen.empty()) { rlen.push_back(0); rmate.push_back(0); rfl.push_back(0); rjlo.push_back(0); rjcut.push_back(0); roff.push_back(0); }
        if (codes.empty()) codes.push_back(4);
        if (flank.empty()) flank.push_back(4);
        if (pileL.empty()) { pileL.assign(4, 0); pileR.assign(4, 0); }
        FbGapBatch B{};
        B.n_gaps = (int)fg.size(); B.gaps = fg.data(); B.n_reads = (int)rlen.size(); B.read_len = rlen.data(); B.read_code_off = roff.data(); B.read_mate = rmate.data();
        B.read_flags = rfl.data(); B.read_jlo = rjlo.data(); B.read_jcut = rjcut.data(); B.n_codes = (int64_t)codes.size(); B.read_codes = codes.data();
        B.n_flank = (int64_t)flank.size(); B.flank_codes = flank.data(); B.n_pile_rows = (int64_t)(pileL.size() / 4); B.pile_left = pileL.data(); B.pile_right = pileR.data();
        if (fb_batch_upload(ctx, &B) != FB_OK) { devErr[d] = std::string("fb_batch_upload: ") + fb_last_error(ctx); fb_ctx_destroy(ctx); return; }
        // The model is needed by the first engine call, not before: the gaps' fibers start at once and build their first requests
        // (nothing in the control logic reads the model before results come back) while the model thread finishes; the lane's
        // first flush waits for it and uploads it.
        double modelWait = 0;
        auto uploadModel = [&]() -> std::string {
            auto w0 = clk::now();
            const bool ok = waitModel();
            modelWait = secs(w0, clk::now());
            if (!ok) return "the model could not be learned";      // (the main thread reports why)
            FbModel fm{};
            fm.max_read_len = model.maxReadLength; fm.err_pos = model.errorPosDist.data(); fm.ins_pos = model.inPosDist.data(); fm.del_pos = model.delPosDist.data();
            for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) fm.err_type[i * 5 + j] = model.errorTypeProbs[i][j];
            fm.n_insert = model.maxInsertSize; fm.insert_pdf = model.insertPdfSmoothed.data();
            fm.insert_min = model.insertThresholdMin; fm.insert_max = model.insertThresholdMax; fm.prob_cutoff = model.gapProbCutOff;
            if (fb_model_upload(ctx, &fm) != FB_OK) return std::string("fb_model_upload: ") + fb_last_error(ctx);
            return std::string();
        };

        auto d1 = clk::now();
        devCtx[d] = secs(d0, d1);
        int inflight = 1024;
        if (const char* e = getenv("FIGBIRD_INFLIGHT")) inflight = std::max(1, atoi(e));
        // worker threads of this lane: the lanes of one GPU alternate (one waits for its kernels while the other replays), so
        // every lane may use the host threads of its GPU
        int nWorkers = std::max(1, hostThreads / std::max(1, nGpus));
        if (laneSpecs[d].kind == 1) nWorkers = std::max(1, nWorkers / 4);
        if (const char* e = getenv("FIGBIRD_LANE_WORKERS")) nWorkers = std::max(1, atoi(e));
        nWorkers = std::max(1, std::min(nWorkers, (int)mine.size()));
        const int cap = std::max(1, (std::min((int)mine.size(), inflight) + nWorkers - 1) / nWorkers);
        LaneQueue q(ctx, nWorkers, uploadModel);
        LaneWork lw; lw.mine = &mine; lw.fills = &fills; lw.results = &results;
        std::vector<std::thread> th;
        std::atomic<long long> cpuNs(0);
        for (int w = 0; w < nWorkers; w++) th.emplace_back([&] {
            laneWorker(q, lw, cap);
            timespec ts; clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts); cpuNs += (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec;
        });
        for (auto& t : th) t.join();
        if (!lw.error.empty()) devErr[d] = lw.error;
        if (!devErr[d].empty()) { fb_ctx_destroy(ctx); return; }      // a CUDA error is sticky: never pool a context that has failed
        devWork[d] = secs(d1, clk::now()); devCpu[d] = cpuNs.load() * 1e-9;
        devCtx[d] += modelWait;      // (t_ctx_upload keeps its meaning: context, batch upload and the wait for the model)
        {   // this run's share of the context's cumulative counters
            FbCounters c1{}; fb_get_counters(ctx, &c1);
            c1.placements_p1 -= ctr0.placements_p1; c1.placements_p2 -= ctr0.placements_p2; c1.base_terms -= ctr0.base_terms; c1.kernel_launches -= ctr0.kernel_launches;
            c1.device_ms -= ctr0.device_ms; c1.h2d_bytes -= ctr0.h2d_bytes; c1.d2h_bytes -= ctr0.d2h_bytes; c1.lane_steps_p1 -= ctr0.lane_steps_p1; c1.lane_steps_p2 -= ctr0.lane_steps_p2;
            c1.device_union_ms -= ctr0.device_union_ms;
            devCtr[d] = c1;
        }
        devTicks[d] = q.ticks(); devEng[d] = q.engineSeconds(); q.dumpTickLog(d);
        releaseCtx(devs[d], laneOf, ctx);
    });
    for (auto& t : devThreads) t.join();
    if (!waitModel()) { printf("%s\n", err.c_str()); return 1; }
    for (auto& e : devErr) if (!e.empty()) { fprintf(stderr, "figbird_b200: %s\n", e.c_str()); return 1; }
    auto t4 = clk::now();

    // ---- outputs (gaps beyond the shorter of gapInfo/stat2 keep an empty entry, like a missing worker line)
    {   // the three output files are independent: written side by side
        bool okGapout = true, okFilled = true;
        std::thread tg([&] { okGapout = writeGapout(a.tmpDir + "gapout.txt", gaps, results); });
        std::thread td([&] {
            // draw.txt = the gaps' texts in the order the reference's workers would have left them (FIGBIRD_DRAW_ORDER=gap: in gap
            // order): sizes are known, so a few threads write disjoint ranges of the file
            const int fd = open((a.tmpDir + "draw.txt").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (fd < 0) return;
            std::vector<int> seq;
            int firstWorker = 0, workers = 1;
            { const char* e = getenv("FIGBIRD_DRAW_ORDER"); if (!(e && !strcmp(e, "gap"))) seq = referenceDrawOrder(a.tmpDir, totGaps, a.numThreads, firstWorker, workers); }
            // mergeFiles (FillGaps.cpp:222-258) appends worker i's file with `out << a.rdbuf() << b.rdbuf()`: inserting an EMPTY a.rdbuf()
            // sets failbit on `out`, b is then dropped, and the merged file stays empty from there on -- so when worker 0 drew nothing
            // and there is more than one worker, the reference's draw.txt is empty.  Reproduced.
            if (workers >= 2 && !seq.empty()) {
                size_t first = 0;
                for (int i = 0; i < firstWorker && i < (int)seq.size(); i++) if (seq[(size_t)i] >= 0 && (size_t)seq[(size_t)i] < results.size()) first += results[(size_t)seq[(size_t)i]].drawText.size();
                if (first == 0) { close(fd); return; }
            }
            {   // a permutation of the gaps that have a result, whatever the dealing said (gap order for anything it left out)
                std::vector<int> keep; std::vector<char> seen(results.size(), 0);
                for (int g : seq) if (g >= 0 && (size_t)g < results.size() && !seen[(size_t)g]) { seen[(size_t)g] = 1; keep.push_back(g); }
                for (size_t g = 0; g < results.size(); g++) if (!seen[g]) keep.push_back((int)g);
                seq.swap(keep);
            }
            std::vector<size_t> off(results.size() + 1, 0);
            for (size_t i = 0; i < results.size(); i++) off[i + 1] = off[i] + results[(size_t)seq[i]].drawText.size();
            const size_t total = off[results.size()];
            int nt = (int)std::max<size_t>(1, std::min<size_t>(6, total >> 24));
            if (const char* e = getenv("FIGBIRD_DRAW_THREADS")) nt = std::max(1, atoi(e));      // (tests force the threaded path on small files)
            auto writeRange = [&](size_t lo, size_t hi) {
                std::string buf; buf.reserve(4u << 20);
                size_t pos = off[lo];
                auto flushBuf = [&] { size_t o2 = 0; while (o2 < buf.size()) { ssize_t k = pwrite(fd, buf.data() + o2, buf.size() - o2, (off_t)(pos + o2)); if (k <= 0) break; o2 += (size_t)k; } pos += buf.size(); buf.clear(); };
                for (size_t i = lo; i < hi; i++) { buf += results[(size_t)seq[i]].drawText; if (buf.size() >= (2u << 20)) flushBuf(); }
                flushBuf();
            };
            if (nt <= 1) writeRange(0, results.size());
            else {
                std::vector<std::thread> th;
                size_t lo = 0;
                for (int t = 0; t < nt; t++) {      // ranges of about equal bytes
                    size_t hi = lo; const size_t want = total * (t + 1) / nt;
                    while (hi < results.size() && off[hi] < want) hi++;
                    if (t == nt - 1) hi = results.size();
                    th.emplace_back(writeRange, lo, hi); lo = hi;
                }
                for (auto& x : th) x.join();
            }
            close(fd);
        });
        okFilled = writeFilledContigs(a.tmpDir, sc, gaps, results, totGaps);
        tg.join(); td.join();
        if (!okGapout) { fprintf(stderr, "cannot write gapout.txt\n"); return 1; }
        if (!okFilled) { fprintf(stderr, "cannot write filledContigs.fa\n"); return 1; }
    }
    auto t5 = clk::now();
    rs.tLoad = secs(t0, t1); rs.tModel = modelSecs; rs.tPrep = secs(t2, t3); rs.tFill = secs(t3, t4); rs.tWrite = secs(t4, t5);
    for (auto& r : results) rs.refPlacements += r.refPlacements;
    for (int d = 0; d < nD; d++) {
        rs.dev.placements_p1 += devCtr[d].placements_p1; rs.dev.placements_p2 += devCtr[d].placements_p2; rs.dev.base_terms += devCtr[d].base_terms;
        rs.dev.kernel_launches += devCtr[d].kernel_launches;
        {   // lanes of one GPU overlap their kernels: the device time of a GPU is the union of its kernel intervals, which the
            // engine tracks per physical device (same value in every lane of that GPU); across GPUs take the maximum
            double u = devCtr[d].device_union_ms;
            for (int e2 = 0; e2 < nD; e2++) if (devs[e2] == devs[d]) u = std::max(u, devCtr[e2].device_union_ms);
            rs.dev.device_ms = std::max(rs.dev.device_ms, u > 0 ? u : devCtr[d].device_ms);
        }
        rs.dev.h2d_bytes += devCtr[d].h2d_bytes; rs.dev.d2h_bytes += devCtr[d].d2h_bytes; rs.dev.lane_steps_p1 += devCtr[d].lane_steps_p1; rs.dev.lane_steps_p2 += devCtr[d].lane_steps_p2; rs.ticks += devTicks[d]; rs.tEngine = std::max(rs.tEngine, devEng[d]); rs.tCopy = std::max(rs.tCopy, devCopy[d]); rs.tCtx = std::max(rs.tCtx, devCtx[d]); rs.tWorkers = std::max(rs.tWorkers, devWork[d]); rs.cpuWorkers += devCpu[d];
    }
    std::string perGpu;      // kernel time (union of the kernel intervals) of every GPU of the run, in FIGBIRD_GPUS order
    {
        std::vector<int> seen;
        for (int d = 0; d < nD; d++) {
            if (std::find(seen.begin(), seen.end(), devs[d]) != seen.end()) continue;
            seen.push_back(devs[d]);
            double u = 0, sum = 0;
            for (int e2 = 0; e2 < nD; e2++) if (devs[e2] == devs[d]) { u = std::max(u, devCtr[e2].device_union_ms); sum += devCtr[e2].device_ms; }
            char b[64]; snprintf(b, sizeof b, "%s%.6f", perGpu.empty() ? "" : ", ", u > 0 ? u : sum);
            perGpu += b;
        }
    }
    if (const char* mp = getenv("FIGBIRD_METRICS")) {
        FILE* mf = fopen(mp, "w");
        if (mf) {
            fprintf(mf, "{\"engine\": \"%s\", \"gaps\": %d, \"gpus\": %d, \"t_load\": %.6f, \"t_model\": %.6f, \"t_prepare\": %.6f, \"t_fill\": %.6f, \"t_write\": %.6f, \"t_engine_calls\": %.6f, \"t_result_copy\": %.6f, \"t_ctx_upload\": %.6f, \"t_workers\": %.6f, \"cpu_workers\": %.6f, "
                        "\"ref_placements_p1\": %lld, \"dev_placements_p1\": %lld, \"dev_placements_p2\": %lld, \"dev_base_terms\": %lld, \"kernel_launches\": %lld, "
                        "\"device_ms\": %.6f, \"h2d_bytes\": %lld, \"d2h_bytes\": %lld, \"ticks\": %lld, \"lane_steps_p1\": %lld, \"lane_steps_p2\": %lld, \"device_ms_per_gpu\": [%s]}\n",
                    fb_engine_name(), nG, nD, rs.tLoad, rs.tModel, rs.tPrep, rs.tFill, rs.tWrite, rs.tEngine, rs.tCopy, rs.tCtx, rs.tWorkers, rs.cpuWorkers, (long long)rs.refPlacements, (long long)rs.dev.placements_p1,
                    (long long)rs.dev.placements_p2, (long long)rs.dev.base_terms, (long long)rs.dev.kernel_launches, rs.dev.device_ms, (long long)rs.dev.h2d_bytes,
                    (long long)rs.dev.d2h_bytes, (long long)rs.ticks, (long long)rs.dev.lane_steps_p1, (long long)rs.dev.lane_steps_p2, perGpu.c_str());
            fclose(mf);
        }
    }
    if (!quiet) printf("Time taken = %d seconds\n", (int)secs(t0, t5));
    if (!quiet) printf("======================================\nIteration %d ends successfully\n======================================\n", a.scriptItr);
    return 0;
}

}  // namespace fb

extern "C" int32_t fb_fillgaps_main(int32_t argc, const char* const* argv) {
    try { return fb::fillgapsMain(argc, argv); }
    catch (const std::exception& e) { fprintf(stderr, "figbird_b200: %s\n", e.what()); return 1; }
}
