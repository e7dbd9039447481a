// jinc_hostmem.cpp -- registry of the caller's host frame buffers and the helper threads of the staging copies.
// See jinc_hostmem.h.  Everything here is host-side bookkeeping around JincResize_GetFrame's frame buffers
// (src/JincResize.cpp:603-630); no pixel is computed here.
#include "jinc_hostmem.h"

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

namespace jinc_hostmem {

namespace {

struct Entry {
    enum State { SEEN, EXTERNAL, REGISTERED, REGISTERING, REFUSED };
    uintptr_t lo = 0, hi = 0;         // the buffer as the caller describes it
    uintptr_t reg_lo = 0, reg_hi = 0; // the page-aligned range registered here
    State state = SEEN;
    int sightings = 0;
    int users = 0;
    bool poisoned = false;
    uint64_t tick = 0;
    double last_use = 0.0; // seconds (steady clock)
};

std::mutex g_mu;
std::map<uintptr_t, Entry*> g_entries; // by lo
size_t g_reg_bytes = 0;
uint64_t g_tick = 0;
long g_registrations = 0;

int g_clients = 0;        // live filters that asked for registration
double g_last_sweep = 0.0;

constexpr size_t kMaxEntries = 4096;
constexpr double kIdleSeconds = 5.0;  // a registration not used for this long is dropped (the host may have retired the buffer)
constexpr double kSweepEvery = 0.25;

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
constexpr uintptr_t kGroupGap = 64 << 10; // planes closer than this belong to one allocation

size_t budget_bytes()
{
    static const size_t b = [] {
        const char* s = getenv("JINCRESIZE_B200_HOSTREG_MB");
        const long mb = s ? atol(s) : 8192;
        return static_cast<size_t>(mb < 0 ? 0 : mb) << 20;
    }();
    return b;
}

uintptr_t page_size()
{
    static const uintptr_t p = [] {
        const long v = sysconf(_SC_PAGESIZE);
        return static_cast<uintptr_t>(v > 0 ? v : 4096);
    }();
    return p;
}

bool is_pinned_host(uintptr_t p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, reinterpret_cast<const void*>(p)) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// g_mu held
void unregister_locked(Entry* e)
{
    if (e->state == Entry::REGISTERED) {
        if (cudaHostUnregister(reinterpret_cast<void*>(e->reg_lo)) != cudaSuccess)
            cudaGetLastError();
        g_reg_bytes -= e->reg_hi - e->reg_lo;
        e->state = Entry::SEEN;
        e->reg_lo = e->reg_hi = 0;
    }
}

// g_mu held: least-recently-used idle registrations go until `need` more bytes fit the budget
bool make_room_locked(size_t need)
{
    if (need > budget_bytes())
        return false;
    while (g_reg_bytes + need > budget_bytes()) {
        Entry* victim = nullptr;
        for (auto& kv : g_entries) {
            Entry* e = kv.second;
            if (e->state == Entry::REGISTERED && e->users == 0 && (!victim || e->tick < victim->tick))
                victim = e;
        }
        if (!victim)
            return false;
        unregister_locked(victim);
        victim->sightings = 0;
    }
    return true;
}

// g_mu held: registrations that have not been used for a while are dropped -- a buffer the host still recycles is seen
// again within a frame time or so, one it has retired (or freed) must not stay page-locked
void sweep_locked(double now)
{
    if (now - g_last_sweep < kSweepEvery)
        return;
    g_last_sweep = now;
    for (auto& kv : g_entries) {
        Entry* e = kv.second;
        if (e->state == Entry::REGISTERED && e->users == 0 && now - e->last_use > kIdleSeconds) {
            unregister_locked(e);
            e->sightings = 1; // registered again the next time it comes back
        }
    }
}

// g_mu held: forget buffers that never came back once the table grows
void prune_locked()
{
    if (g_entries.size() < kMaxEntries)
        return;
    std::vector<std::pair<uint64_t, uintptr_t>> idle;
    for (auto& kv : g_entries)
        if (kv.second->users == 0 && (kv.second->state == Entry::SEEN || kv.second->state == Entry::REFUSED))
            idle.emplace_back(kv.second->tick, kv.first);
    std::sort(idle.begin(), idle.end());
    for (size_t i = 0; i < idle.size() / 2; ++i) {
        auto it = g_entries.find(idle[i].second);
        delete it->second;
        g_entries.erase(it);
    }
}

// registers e's range; called WITHOUT g_mu (page-locking a large frame buffer takes milliseconds), e->state is REGISTERING
bool do_register(Entry* e)
{
    // whole pages INSIDE the buffer only: the partial pages at its ends are shared with whatever the host keeps next to it
    const uintptr_t ps = page_size();
    const uintptr_t lo = (e->lo + ps - 1) & ~(ps - 1), hi = e->hi & ~(ps - 1);
    if (hi <= lo || hi - lo < 16 * ps)
        return false; // not worth a registration
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!make_room_locked(hi - lo))
            return false;
        g_reg_bytes += hi - lo; // reserved
    }
    cudaError_t err = cudaHostRegister(reinterpret_cast<void*>(lo), hi - lo, cudaHostRegisterPortable);
    if (err == cudaErrorHostMemoryAlreadyRegistered) {
        // a stale registration of ours overlaps (the host re-cut its memory): drop the idle ones and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lk(g_mu);
            for (auto& kv : g_entries) {
                Entry* o = kv.second;
                if (o != e && o->state == Entry::REGISTERED && o->users == 0 && o->reg_lo < hi && lo < o->reg_hi) {
                    unregister_locked(o);
                    o->sightings = 0;
                }
            }
        }
        err = cudaHostRegister(reinterpret_cast<void*>(lo), hi - lo, cudaHostRegisterPortable);
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (err != cudaSuccess) {
        cudaGetLastError();
        g_reg_bytes -= hi - lo;
        if (getenv("JINCRESIZE_B200_DEBUG"))
            fprintf(stderr, "jinc_hostmem: cudaHostRegister(%p, %zu) failed: %s\n", reinterpret_cast<void*>(lo), static_cast<size_t>(hi - lo),
                    cudaGetErrorString(err));
        return false;
    }
    e->reg_lo = lo;
    e->reg_hi = hi;
    ++g_registrations;
    return true;
}

} // namespace

// ------------------------------------------------------------------------------------------ helper threads

namespace {

class CopyPool {
public:
    // the helpers of the staging copies (never destroyed: helper threads may outlive static destructors)
    static CopyPool& get()
    {
        static CopyPool* p = [] {
            const char* s = getenv("JINCRESIZE_B200_COPY_THREADS");
            const int hw = static_cast<int>(std::thread::hardware_concurrency());
            return new CopyPool(s ? atoi(s) : std::min(3, std::max(0, hw / 4 - 1)));
        }();
        return *p;
    }
    // one thread of its own for registrations: page-locking hundreds of megabytes must not sit in front of a staging copy
    static CopyPool& registrar()
    {
        static CopyPool* p = new CopyPool(1);
        return *p;
    }
    int helpers() const { return static_cast<int>(threads_.size()); }
    void run(std::function<void()> fn)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(std::move(fn));
        }
        cv_.notify_one();
    }

private:
    explicit CopyPool(int n)
    {
        n = std::max(0, std::min(n, 16));
        for (int i = 0; i < n; ++i) {
            threads_.emplace_back([this] {
                for (;;) {
                    std::function<void()> fn;
                    {
                        std::unique_lock<std::mutex> lk(mu_);
                        cv_.wait(lk, [this] { return !q_.empty(); });
                        fn = std::move(q_.front());
                        q_.pop_front();
                    }
                    fn();
                }
            });
            threads_.back().detach();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> threads_;
};

} // namespace

bool acquire(const Range* ranges, int n, bool may_register, int device, Pin* pin)
{
    pin->n = 0;
    if (n < 1 || n > JINC_MAX_PLANES)
        return false;
    // planes of one frame buffer form one range
    Range sorted[JINC_MAX_PLANES];
    for (int i = 0; i < n; ++i) {
        if (!ranges[i].lo || ranges[i].hi <= ranges[i].lo)
            return false;
        sorted[i] = ranges[i];
    }
    std::sort(sorted, sorted + n, [](const Range& a, const Range& b) { return a.lo < b.lo; });
    uintptr_t glo[JINC_MAX_PLANES], ghi[JINC_MAX_PLANES];
    int ng = 0;
    for (int i = 0; i < n; ++i) {
        const uintptr_t lo = reinterpret_cast<uintptr_t>(sorted[i].lo), hi = reinterpret_cast<uintptr_t>(sorted[i].hi);
        if (ng > 0 && lo <= ghi[ng - 1] + kGroupGap) {
            ghi[ng - 1] = std::max(ghi[ng - 1], hi);
        } else {
            glo[ng] = lo;
            ghi[ng] = hi;
            ++ng;
        }
    }

    Entry* got[JINC_MAX_PLANES];
    const double now = now_s();
    std::unique_lock<std::mutex> lk(g_mu);
    sweep_locked(now);
    for (int g = 0; g < ng; ++g) {
        Entry* e = nullptr;
        auto it = g_entries.find(glo[g]);
        if (it != g_entries.end()) {
            e = it->second;
            if (e->hi != ghi[g]) { // another buffer now lives at this address
                if (e->users > 0 || e->state == Entry::REGISTERING)
                    return false;
                unregister_locked(e);
                delete e;
                g_entries.erase(it);
                e = nullptr;
            }
        }
        if (!e) {
            prune_locked();
            e = new Entry();
            e->lo = glo[g];
            e->hi = ghi[g];
            e->state = (is_pinned_host(glo[g]) && is_pinned_host(ghi[g] - 1)) ? Entry::EXTERNAL : Entry::SEEN;
            g_entries.emplace(glo[g], e);
        }
        ++e->sightings;
        e->tick = ++g_tick;
        e->last_use = now;
        if (e->state == Entry::SEEN && may_register && e->sightings >= 2 && !e->poisoned) {
            // Page-locking a large frame buffer takes tens of milliseconds: it happens on a helper thread while this
            // frame is staged like the first one, and the buffer is used directly from its next return on.
            e->state = Entry::REGISTERING;
            auto task = [e, device] {
                cudaSetDevice(device);
                const bool ok = do_register(e);
                std::lock_guard<std::mutex> lk2(g_mu);
                e->state = ok ? Entry::REGISTERED : Entry::REFUSED;
                if (ok && g_clients == 0) { // the last registering filter went away meanwhile
                    unregister_locked(e);
                    e->sightings = 0;
                }
            };
            CopyPool::registrar().run(task);
        }
        if (e->state != Entry::EXTERNAL && e->state != Entry::REGISTERED)
            return false;
        got[g] = e;
    }
    for (int g = 0; g < ng; ++g) {
        Entry* e = got[g];
        ++e->users;
        pin->entry[g] = e;
        pin->lo[g] = e->lo;
        pin->hi[g] = e->hi;
        pin->dlo[g] = e->state == Entry::REGISTERED ? e->reg_lo : e->lo;
        pin->dhi[g] = e->state == Entry::REGISTERED ? e->reg_hi : e->hi;
    }
    pin->n = ng;
    return true;
}

void direct_part(const Pin& pin, const void* p, size_t n, size_t* a, size_t* b)
{
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p), hi = lo + n;
    *a = *b = 0;
    for (int g = 0; g < pin.n; ++g) {
        if (lo < pin.lo[g] || hi > pin.hi[g])
            continue;
        const uintptr_t x = std::min(std::max(pin.dlo[g], lo), hi), y = std::min(std::max(pin.dhi[g], lo), hi);
        if (y > x) {
            *a = x - lo;
            *b = y - lo;
        }
        return;
    }
}

void client_add()
{
    std::lock_guard<std::mutex> lk(g_mu);
    ++g_clients;
}

void client_remove()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (--g_clients > 0)
        return;
    // the last filter that registers caller memory is gone: nothing stays page-locked behind the host's back
    for (auto it = g_entries.begin(); it != g_entries.end();) {
        Entry* e = it->second;
        if (e->users == 0 && e->state != Entry::REGISTERING) {
            unregister_locked(e);
            delete e;
            it = g_entries.erase(it);
        } else {
            ++it;
        }
    }
}

void release(Pin* pin)
{
    if (pin->n == 0)
        return;
    const double now = now_s();
    std::lock_guard<std::mutex> lk(g_mu);
    for (int g = 0; g < pin->n; ++g) {
        Entry* e = static_cast<Entry*>(pin->entry[g]);
        e->last_use = now;
        if (--e->users == 0 && e->poisoned) {
            unregister_locked(e);
            e->state = Entry::REFUSED;
        }
    }
    pin->n = 0;
}

bool registered_here(const Pin* pin)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (int g = 0; g < pin->n; ++g)
        if (static_cast<const Entry*>(pin->entry[g])->state == Entry::REGISTERED)
            return true;
    return false;
}

void distrust(Pin* pin)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (int g = 0; g < pin->n; ++g) {
        Entry* e = static_cast<Entry*>(pin->entry[g]);
        if (e->state == Entry::REGISTERED)
            e->poisoned = true;
    }
}

size_t registered_bytes()
{
    std::lock_guard<std::mutex> lk(g_mu);
    return g_reg_bytes;
}

long registrations()
{
    std::lock_guard<std::mutex> lk(g_mu);
    return g_registrations;
}

// ------------------------------------------------------------------------------------------ staging copies

namespace {

void copy_rows_serial(unsigned char* dst, size_t dst_pitch, const unsigned char* src, ptrdiff_t src_pitch, size_t row_bytes, int rows)
{
    if (static_cast<ptrdiff_t>(dst_pitch) == src_pitch && dst_pitch == row_bytes) {
        memcpy(dst, src, row_bytes * static_cast<size_t>(rows));
        return;
    }
    for (int y = 0; y < rows; ++y)
        memcpy(dst + static_cast<size_t>(y) * dst_pitch, src + static_cast<ptrdiff_t>(y) * src_pitch, row_bytes);
}

} // namespace

void copy_rows(unsigned char* dst, size_t dst_pitch, const unsigned char* src, ptrdiff_t src_pitch, size_t row_bytes, int rows)
{
    const size_t total = row_bytes * static_cast<size_t>(rows);
    constexpr size_t kPart = static_cast<size_t>(1) << 20;
    CopyPool& pool = CopyPool::get();
    const int parts = static_cast<int>(std::min<size_t>(pool.helpers() + 1, total / kPart));
    if (parts < 2 || rows < parts) {
        copy_rows_serial(dst, dst_pitch, src, src_pitch, row_bytes, rows);
        return;
    }
    struct Latch {
        std::mutex mu;
        std::condition_variable cv;
        int left;
    } latch;
    latch.left = parts - 1;
    const int per = (rows + parts - 1) / parts;
    for (int p = 1; p < parts; ++p) {
        const int r0 = p * per, r1 = std::min(rows, r0 + per);
        pool.run([=, &latch] {
            if (r1 > r0)
                copy_rows_serial(dst + static_cast<size_t>(r0) * dst_pitch, dst_pitch, src + static_cast<ptrdiff_t>(r0) * src_pitch, src_pitch,
                                 row_bytes, r1 - r0);
            std::lock_guard<std::mutex> lk(latch.mu);
            if (--latch.left == 0)
                latch.cv.notify_one();
        });
    }
    copy_rows_serial(dst, dst_pitch, src, src_pitch, row_bytes, std::min(rows, per));
    std::unique_lock<std::mutex> lk(latch.mu);
    latch.cv.wait(lk, [&] { return latch.left == 0; });
}

} // namespace jinc_hostmem
