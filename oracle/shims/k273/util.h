/* ORACLE shim for k273/util.h: wall clock, LockedQueue, WorkerInterface / WorkerThread
 * (used by src/cpp/supervisor.h:29-70, supervisor.cpp:79-99,196-245). */
#pragma once
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace K273 {

inline double get_time() {
    using namespace std::chrono;
    return duration_cast<duration<double>>(steady_clock::now().time_since_epoch()).count();
}

template <typename T>
class LockedQueue {
public:
    void push(T v) { std::lock_guard<std::mutex> lk(m); q.push_back(v); }
    T pop() { std::lock_guard<std::mutex> lk(m); T v = q.front(); q.pop_front(); return v; }
    bool empty() { std::lock_guard<std::mutex> lk(m); return q.empty(); }
private:
    std::mutex m;
    std::deque<T> q;
};

class WorkerThread;

class WorkerInterface {
public:
    WorkerInterface() : thread_self(nullptr) {}
    virtual ~WorkerInterface() {}
    virtual void doWork() = 0;
    void setThread(WorkerThread* t) { thread_self = t; }
    WorkerThread* getThread() { return thread_self; }
protected:
    WorkerThread* thread_self;
};

class WorkerThread {
public:
    explicit WorkerThread(WorkerInterface* w) : worker(w), prompted(false) { w->setThread(this); }
    void spawn() {
        th = std::thread([this]() {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [this]() { return prompted; });
            lk.unlock();
            worker->doWork();
        });
        th.detach();
    }
    void startPolling() {}
    void promptWorker() {
        { std::lock_guard<std::mutex> lk(m); prompted = true; }
        cv.notify_all();
    }
    void done() {}
private:
    WorkerInterface* worker;
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    bool prompted;
};

}  // namespace K273
