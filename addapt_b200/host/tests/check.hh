// Minimal test harness for the host facade (Catch, which the reference's tests use, is not in this image).
#pragma once
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

struct TestCase { const char* name; bool needs_gpu; std::function<void()> body; };
std::vector<TestCase>& registry();
extern int g_checks, g_failures;

struct Registrar {
  Registrar(const char* name, bool gpu, std::function<void()> body) { registry().push_back({name, gpu, body}); }
};

#define TEST_CASE(ident, name) static void ident(); static Registrar reg_##ident(name, false, ident); static void ident()
#define GPU_TEST_CASE(ident, name) static void ident(); static Registrar reg_##ident(name, true, ident); static void ident()

#define CHECK(...)                                                                         \
  do {                                                                                      \
    g_checks++;                                                                             \
    bool ok_ = false;                                                                       \
    try { ok_ = static_cast<bool>(__VA_ARGS__); } catch (...) { ok_ = false; }                     \
    if (!ok_) { g_failures++; std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #__VA_ARGS__); } \
  } while (0)
#define REQUIRE(...) CHECK(__VA_ARGS__)
#define CHECK_THROWS(...)                                                                  \
  do {                                                                                      \
    g_checks++;                                                                             \
    bool threw_ = false;                                                                    \
    try { (void)(__VA_ARGS__); } catch (...) { threw_ = true; }                                    \
    if (!threw_) { g_failures++; std::printf("  FAILED %s:%d: no exception from %s\n", __FILE__, __LINE__, #__VA_ARGS__); } \
  } while (0)

inline bool approx(double a, double b, double rel = 1e-9) { return std::fabs(a - b) <= rel * std::fmax(1.0, std::fmax(std::fabs(a), std::fabs(b))); }
