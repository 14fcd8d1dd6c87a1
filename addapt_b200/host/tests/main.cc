// run_host_tests [--gpu]: without the flag only the CPU-side cases run (model, sampling rules, score-function
// plumbing with fake fold engines, config parsing); with it, also the cases that fold on the B200 engine.
#include <cstring>
#include <string>

#include "check.hh"

std::vector<TestCase>& registry() { static std::vector<TestCase> r; return r; }
int g_checks = 0, g_failures = 0;

int main(int argc, char** argv) {
  bool gpu = false;
  for (int k = 1; k < argc; k++) gpu |= std::strcmp(argv[k], "--gpu") == 0;
  int ran = 0;
  for (auto& t : registry()) {
    if (t.needs_gpu != gpu) continue;
    std::printf("[%s] %s\n", t.needs_gpu ? "gpu" : "cpu", t.name);
    const int before = g_failures;
    try { t.body(); }
    catch (std::string& e) { g_failures++; std::printf("  EXCEPTION: %s\n", e.c_str()); }
    catch (const char* e) { g_failures++; std::printf("  EXCEPTION: %s\n", e); }
    catch (std::exception& e) { g_failures++; std::printf("  EXCEPTION: %s\n", e.what()); }
    if (g_failures != before) std::printf("  ^^^ case failed\n");
    ran++;
  }
  std::printf("%d cases, %d checks, %d failures\n", ran, g_checks, g_failures);
  return g_failures ? 1 : 0;
}
