/* ORACLE shim for k273/exception.h */
#pragma once
#include <cstdio>
#include <cstdlib>
#include <string>

namespace K273 {
class Exception {
public:
    explicit Exception(const std::string& m) : msg(m) {}
    const std::string& getMessage() const { return msg; }
private:
    std::string msg;
};
}  // namespace K273

#define ASSERT(cond)                                                                     \
    do { if (!(cond)) { std::fprintf(stderr, "ASSERT failed: %s at %s:%d\n", #cond,       \
                                     __FILE__, __LINE__); std::abort(); } } while (0)
#define ASSERT_MSG(cond, msg)                                                            \
    do { if (!(cond)) { std::fprintf(stderr, "ASSERT failed: %s (%s) at %s:%d\n", #cond,  \
                                     msg, __FILE__, __LINE__); std::abort(); } } while (0)
