// jinc_hostmem.h -- what the frame pipeline knows about the caller's host buffers (not part of the public ABI).
//
// The reference's GetFrame works in place on AviSynth's frame buffers (src/JincResize.cpp:603-630).  A GPU path has to
// move them over PCIe, and the DMA engines only run at full rate -- and only overlap with the caller -- from page-locked
// memory.  A host recycles its frame buffers (AviSynth's frame cache, the mini-host's pool), so the pipeline remembers
// every buffer it has seen: buffers the caller allocated page-locked are used directly; pageable buffers are staged
// through the slot's pinned mirror the first time and registered (cudaHostRegister) when they come back, after which
// they are used directly too.  Registered memory is capped and evicted least-recently-used; a registration that has
// not been used for five seconds is dropped, and everything is unregistered when the last registering filter goes.
//
// Registration is OPT-IN (JINC_FILTER_HOST_REGISTER) because it rests on a promise only the host can make: a buffer
// that was registered must not be freed while it still is.  A stale registration sends DMA to the buffer's former
// physical pages and makes unrelated CUDA calls that touch the re-used address range fail.  The pipeline therefore
// checks every transfer through a registration it made (arrival sentinels in destination planes, probe words read back
// from source planes) and drops the registration -- redoing the frame through the staged path -- on a miss.
#ifndef JINC_HOSTMEM_H
#define JINC_HOSTMEM_H

#include <cstddef>
#include <cstdint>

#include "jinc_b200.h"

namespace jinc_hostmem {

struct Range {
    const unsigned char* lo;
    const unsigned char* hi; // one past the last byte
};

// Ranges held against eviction between acquire() and release(): the buffers (planes closer than 64 KB form one), and of
// each the part that is page-locked -- all of it for memory the caller allocated page-locked, the pages that lie FULLY
// INSIDE the buffer for a registration made here (memory next to the caller's buffer is never locked: a CUDA transfer
// that straddles the edge of a registered range is an error, and those bytes are not ours to lock).
struct Pin {
    int n = 0;
    void* entry[JINC_MAX_PLANES] = {};
    uintptr_t lo[JINC_MAX_PLANES] = {}, hi[JINC_MAX_PLANES] = {};   // the buffer
    uintptr_t dlo[JINC_MAX_PLANES] = {}, dhi[JINC_MAX_PLANES] = {}; // its page-locked part
};

// The part [*a, *b) (offsets into [0, n)) of host range [p, p + n) that DMA may address directly; *a == *b: none.
void direct_part(const Pin& pin, const void* p, size_t n, size_t* a, size_t* b);

// True when every range is page-locked (allocated so by the caller, or registered here); the ranges are then pinned
// against eviction until release().  False: stage the frame.  With `may_register` a pageable buffer that has been seen
// before is handed to a helper thread for registration (on CUDA device `device`) and used directly once that is done.
bool acquire(const Range* ranges, int n, bool may_register, int device, Pin* pin);
void release(Pin* pin);
// filters that may register caller memory; when the last one goes every registration is dropped
void client_add();
void client_remove();
// true when any of the pinned ranges is page-locked by a registration made here (as opposed to by the caller)
bool registered_here(const Pin* pin);
// A transfer through ranges this module registered did not arrive (the host freed and re-mapped the memory behind the
// registration): forget those registrations and never register the addresses again.  Call before release().
void distrust(Pin* pin);
// statistics: bytes currently registered here, number of cudaHostRegister calls so far
size_t registered_bytes();
long registrations();

// rows of one plane, host to host; large planes are split over a small pool of helper threads
void copy_rows(unsigned char* dst, size_t dst_pitch, const unsigned char* src, ptrdiff_t src_pitch, size_t row_bytes, int rows);

} // namespace jinc_hostmem

#endif
