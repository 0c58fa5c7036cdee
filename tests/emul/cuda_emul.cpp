// TEST INFRASTRUCTURE ONLY — see cuda_emul.h.
#include "cuda_emul.h"

#if !defined(__x86_64__)
#error "the CTA emulator's fibre switch is written for x86-64 (the build/test container)"
#endif

// Minimal cooperative context switch (System V x86-64): saves the callee-saved
// registers on the current stack, swaps stack pointers, restores and returns.
asm(R"(
    .text
    .globl pm_emul_swap
    .type pm_emul_swap, @function
pm_emul_swap:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size pm_emul_swap, .-pm_emul_swap
)");

namespace pm_emul {
thread_local Block* g_blk = nullptr;
static const size_t kStack = 192 * 1024;

static void trampoline() {
    Block* b = g_blk;
    (*b->body)();
    b->fibers[b->cur].done = true;
    b->progress++;
    pm_emul_swap(&b->fibers[b->cur].sp, b->sched_sp);
    abort();  // a finished fibre is never resumed
}

void run_block(int block_id, int nblocks, int nthreads, size_t smem_bytes, const std::function<void()>& body) {
    Block blk;
    blk.nthreads = nthreads;
    blk.block_id = block_id;
    blk.nblocks = nblocks;
    blk.body = &body;
    blk.fibers.resize(nthreads);
    blk.warps.resize((nthreads + 31) / 32);
    blk.wgen.assign(nthreads, 0);
    blk.tgen.assign(nthreads, 0);
    blk.smem = (unsigned char*)aligned_alloc(128, (smem_bytes + 127) / 128 * 128 + 128);
    blk.smem_bytes = smem_bytes;
    memset(blk.smem, 0xCD, smem_bytes);  // smem is uninitialised on the GPU too
    Block* saved = g_blk;
    g_blk = &blk;
    for (int t = 0; t < nthreads; t++) {
        Fiber& f = blk.fibers[t];
        f.stack = (char*)aligned_alloc(64, kStack);
        uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;              // fake return address of trampoline (keeps rsp % 16 == 8 at entry)
        *--sp = (void*)trampoline;    // `ret` target of the first switch
        for (int i = 0; i < 6; i++) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
        f.sp = sp;
    }
    int alive = nthreads;
    while (alive > 0) {
        unsigned long long before = blk.progress;
        for (int t = 0; t < nthreads; t++) {
            if (blk.fibers[t].done) continue;
            blk.cur = t;
            pm_emul_swap(&blk.sched_sp, blk.fibers[t].sp);
            if (blk.fibers[t].done) alive--;
        }
        if (alive > 0 && blk.progress == before) {
            fprintf(stderr, "pm_emul: deadlock in block %d (%d threads alive, barrier arrived %d)\n",
                    block_id, alive, blk.bar_arrived);
            abort();
        }
    }
    for (auto& f : blk.fibers) free(f.stack);
    free(blk.smem);
    g_blk = saved;
}
}  // namespace pm_emul
