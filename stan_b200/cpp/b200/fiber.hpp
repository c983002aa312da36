// b200/fiber.hpp -- stackful coroutines for the batched multi-chain driver (b200/batched_nuts.hpp).
// No dependency on the reference or on CUDA: tests/test_fiber_host.py builds and exercises it on its own.
#ifndef B200_FIBER_HPP
#define B200_FIBER_HPP

#include <sys/mman.h>

#include <cstddef>
#include <cstdint>
#include <cxxabi.h>
#include <functional>
#include <stdexcept>
#include <utility>

namespace b200 {

// ---------------------------------------------------------------------------------------------
// Fibers: stackful coroutines, switched in user space (x86-64 System V: the callee-saved registers and the
// stack pointer; no signal-mask system call as swapcontext makes).  A fiber is resumed by the scheduler and
// runs until it yields or its body returns.  The C++ exception-handling globals of the thread (the stack of
// caught exceptions) are swapped with the stacks, so a chain suspended inside a try block cannot see another
// chain's exceptions.
// ---------------------------------------------------------------------------------------------
#if !defined(__x86_64__)
#error "b200/batched_nuts.hpp: the fiber switch is written for x86-64"
#endif
extern "C" void b200_fiber_switch(void** save_sp, void* load_sp);
asm(R"(
.pushsection .text
.weak b200_fiber_switch
.type b200_fiber_switch,@function
b200_fiber_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size b200_fiber_switch,.-b200_fiber_switch
.popsection
)");

class fiber {
 public:
  static constexpr size_t kStackBytes = 512 * 1024;   // build_tree recursion <= max_depth frames of a few hundred bytes
  explicit fiber(std::function<void()> body) : body_(std::move(body)) {
    const size_t page = 4096;
    map_bytes_ = kStackBytes + page;
    map_ = ::mmap(nullptr, map_bytes_, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
    if (map_ == MAP_FAILED)
      throw std::runtime_error("fiber: mmap of the chain stack failed");
    ::mprotect(map_, page, PROT_NONE);                // guard page below the stack
    char* top = static_cast<char*>(map_) + map_bytes_;
    void** sp = reinterpret_cast<void**>(reinterpret_cast<uintptr_t>(top) & ~uintptr_t(15));
    *--sp = nullptr;                                  // return address of the entry function (it never returns)
    *--sp = reinterpret_cast<void*>(&fiber::entry);   // popped by `ret` in b200_fiber_switch
    for (int i = 0; i < 6; ++i)
      *--sp = nullptr;                                // rbp rbx r12 r13 r14 r15
    unsigned int* csr = reinterpret_cast<unsigned int*>(--sp);   // mxcsr | x87 control word: the current thread's
    unsigned int mx = 0;
    unsigned short cw = 0;
    asm volatile("stmxcsr %0" : "=m"(mx));
    asm volatile("fnstcw %0" : "=m"(cw));
    csr[0] = mx;
    csr[1] = cw;
    sp_ = sp;
  }
  fiber(const fiber&) = delete;
  fiber& operator=(const fiber&) = delete;
  ~fiber() {
    if (map_ != MAP_FAILED)
      ::munmap(map_, map_bytes_);
  }
  bool done() const { return done_; }
  // scheduler side: run the fiber until it yields or finishes
  void resume() {
    fiber* prev = current();
    current() = this;
    swap_eh_globals();
    b200_fiber_switch(&sched_sp_, sp_);
    swap_eh_globals();
    current() = prev;
  }
  // fiber side: back to the scheduler
  static void yield() {
    fiber* f = current();
    b200_fiber_switch(&f->sp_, f->sched_sp_);
  }
  static fiber*& current() {
    static thread_local fiber* cur = nullptr;
    return cur;
  }

 private:
  static void entry() {
    fiber* f = current();
    try {
      f->body_();
    } catch (...) {   // nothing may unwind past the bottom of a fiber stack
    }
    f->done_ = true;
    for (;;)
      yield();
  }
  // __cxa_eh_globals = { void* caughtExceptions; unsigned int uncaughtExceptions; } (Itanium C++ ABI)
  void swap_eh_globals() {
    struct eh_globals {
      void* caught;
      unsigned int uncaught;
    };
    eh_globals* g = reinterpret_cast<eh_globals*>(abi::__cxa_get_globals());
    std::swap(g->caught, eh_caught_);
    std::swap(g->uncaught, eh_uncaught_);
  }
  std::function<void()> body_;
  void* map_ = MAP_FAILED;
  size_t map_bytes_ = 0;
  void* sp_ = nullptr;
  void* sched_sp_ = nullptr;
  bool done_ = false;
  void* eh_caught_ = nullptr;
  unsigned int eh_uncaught_ = 0;
};

}  // namespace b200

#endif
