/* ORACLE shim for k273's C-level greenlets (greenlet/greenlet.h), built on ucontext.
 * Surface used by src/cpp/scheduler.{h,cpp}: greenlet_t, greenlet_current, greenlet_switch_to,
 * greenlet_isdead, createGreenlet(callable[, parent]). A greenlet that returns from its function
 * is dead and control passes to its parent. State is per thread (worker mode runs one scheduler
 * per thread). */
#pragma once
#include <functional>

struct greenlet_t;

greenlet_t* greenlet_current();
void greenlet_switch_to(greenlet_t* g);
bool greenlet_isdead(greenlet_t* g);
greenlet_t* gzref_create_greenlet(std::function<void()> fn, greenlet_t* parent);

template <typename Callable>
greenlet_t* createGreenlet(Callable f, greenlet_t* parent = nullptr) {
    return gzref_create_greenlet(std::function<void()>([f]() mutable { f(); }), parent);
}
