// memory_arena member functions for the g++ build of the reference (test infrastructure, oracle/_ref).
// The reference implements them with VirtualAlloc reserve / commit (src/core/memory.cpp, Windows only); this is the same bump
// allocator over one anonymous mmap: pages are reserved up front and committed by the kernel on first touch.
#include "pch.h"
#include "core/memory.h"
#include "core/math.h"
#include <sys/mman.h>

void memory_arena::initialize(uint64 minimumBlockSize_, uint64 reserveSize_)
{
    reset(true);
    void* p = mmap(nullptr, reserveSize_, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    memory = (p == MAP_FAILED) ? nullptr : (uint8*)p;
    pageSize = 4096;
    sizeLeftTotal = reserveSize_;
    minimumBlockSize = minimumBlockSize_;
    reserveSize = reserveSize_;
}

void memory_arena::ensureFreeSize(uint64 size) { mutex.lock(); ensureFreeSizeInternal(size); mutex.unlock(); }

void memory_arena::ensureFreeSizeInternal(uint64 size)
{
    if (sizeLeftCurrent < size)
    {
        uint64 allocationSize = size > minimumBlockSize ? size : minimumBlockSize;
        allocationSize = pageSize * ((allocationSize + pageSize - 1) / pageSize);
        sizeLeftTotal += allocationSize; sizeLeftCurrent += allocationSize; committedMemory += allocationSize;    // commit = first touch
    }
}

void* memory_arena::allocate(uint64 size, uint64 alignment, bool clearToZero)
{
    if (size == 0) return 0;
    mutex.lock();
    const uint64 mask = alignment - 1, misalignment = current & mask, adjustment = (misalignment == 0) ? 0 : (alignment - misalignment);
    current += adjustment; sizeLeftCurrent -= adjustment; sizeLeftTotal -= adjustment;
    ASSERT(sizeLeftTotal >= size);
    ensureFreeSizeInternal(size);
    uint8* result = memory + current;
    current += size; sizeLeftCurrent -= size; sizeLeftTotal -= size;
    mutex.unlock();
    if (clearToZero) memset(result, 0, size);
    return result;
}

void* memory_arena::getCurrent(uint64 alignment) { return memory + alignTo(current, alignment); }
void memory_arena::setCurrentTo(void* ptr) { current = (uint8*)ptr - memory; sizeLeftCurrent = committedMemory - current; sizeLeftTotal = reserveSize - current; }
void memory_arena::reset(bool freeMemory)
{
    if (memory && freeMemory) { munmap(memory, reserveSize); memory = 0; committedMemory = 0; }
    resetToMarker(memory_marker{ 0 });
}
memory_marker memory_arena::getMarker() { return { current }; }
void memory_arena::resetToMarker(memory_marker marker) { current = marker.before; sizeLeftCurrent = committedMemory - current; sizeLeftTotal = reserveSize - current; }
