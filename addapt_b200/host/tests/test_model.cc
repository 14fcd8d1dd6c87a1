// The reference's tests/test_model.cc cases, run against the facade's Device / Aptamer / Context.
#include "addapt/model.hh"
#include "check.hh"

using namespace addapt;

TEST_CASE(model_device, "Device: sequence, indexing, macrostates, copy/assign (reference tests/test_model.cc:8-41)") {
  Device dummy("ACGU");
  CHECK(dummy.seq() == "ACGU");
  CHECK(dummy.len() == 4);
  const char* fwd = "ACGU";
  for (int k = 0; k < 4; k++) CHECK(dummy.seq(k) == fwd[k]);
  CHECK_THROWS(dummy.seq(4));
  for (int k = 1; k <= 4; k++) CHECK(dummy.seq(-k) == fwd[4 - k]);
  CHECK_THROWS(dummy.seq(-5));

  dummy.add_macrostate("a", "....");
  dummy.add_macrostate("b", "(())");
  CHECK(dummy.macrostate("a") == "....");
  CHECK(dummy.macrostate("b") == "(())");
  CHECK_THROWS(dummy.macrostate("missing"));
  CHECK_THROWS(dummy.add_macrostate("short", "..."));

  DevicePtr twin = dummy.copy();
  CHECK(twin->seq() == "ACGU");
  CHECK(twin->macrostate("a") == "....");
  CHECK(twin->macrostate("b") == "(())");

  Device other("nnnn");
  CHECK(other.seq() == "nnnn");
  other.assign(twin);
  CHECK(other.seq() == "ACGU");
  CHECK(other.macrostate("b") == "(())");
}

TEST_CASE(model_context, "Device with a context (reference tests/test_model.cc:43-89)") {
  Device dummy("C");
  dummy.add_macrostate("bp", "x");
  CHECK(dummy.context()->before() == "");
  CHECK(dummy.context()->after() == "");

  dummy.context(make_shared<Context>("A", "GU"));
  CHECK(dummy.len() == 4);
  CHECK(dummy.seq() == "ACGU");
  CHECK(dummy.seq(0) == 'A');
  CHECK(dummy.seq(3) == 'U');
  CHECK(dummy.raw_len() == 1);
  CHECK(dummy.raw_seq() == "C");
  CHECK(dummy.raw_seq(0) == 'C');
  CHECK(dummy.macrostate("bp") == ".x..");
  CHECK(dummy.context()->before() == "A");
  CHECK(dummy.context()->after() == "GU");
  for (auto ms : dummy.macrostates()) { CHECK(ms.first == "bp"); CHECK(ms.second == ".x.."); }

  dummy.remove_context();
  CHECK(dummy.len() == 1);
  CHECK(dummy.seq() == "C");
  CHECK(dummy.macrostate("bp") == "x");
  for (auto ms : dummy.macrostates()) { CHECK(ms.first == "bp"); CHECK(ms.second == "x"); }
}

TEST_CASE(model_mutate, "Device::mutate (reference tests/test_model.cc:91-112)") {
  {
    Device d("AAAA");
    const char* want[] = {"UAAA", "UUAA", "UUUA", "UUUU"};
    for (int k = 0; k < 4; k++) { d.mutate(k, 'U'); CHECK(d.seq() == want[k]); }
  }
  {
    Device d("AAAA");
    const char* want[] = {"AAAU", "AAUU", "AUUU", "UUUU"};
    for (int k = 1; k <= 4; k++) { d.mutate(-k, 'U'); CHECK(d.seq() == want[k - 1]); }
  }
  Device d("AAAA");
  CHECK_THROWS(d.mutate(4, 'U'));
  CHECK_THROWS(d.mutate(-5, 'U'));
}

TEST_CASE(model_aptamer, "Aptamer accessors (reference tests/test_model.cc:114-123)") {
  Aptamer theo("GAUACCAGCCGAAAGGCCCUUGGCAGC", "(...((.(((....)))....))...)", 0.320);
  CHECK(theo.seq() == "GAUACCAGCCGAAAGGCCCUUGGCAGC");
  CHECK(theo.fold() == "(...((.(((....)))....))...)");
  CHECK(theo.affinity() == 0.320);
}

TEST_CASE(utils_index, "normalize_index / normalize_range (reference src/utils.cc:5-40)") {
  CHECK(normalize_index("ACGU", 0, IndexEnum::ITEM) == 0);
  CHECK(normalize_index("ACGU", -1, IndexEnum::ITEM) == 3);
  CHECK(normalize_index("ACGU", 4, IndexEnum::BETWEEN) == 4);
  CHECK(normalize_index("ACGU", -1, IndexEnum::BETWEEN) == 4);
  CHECK_THROWS(normalize_index("ACGU", 4, IndexEnum::ITEM));
  CHECK_THROWS(normalize_index("ACGU", 5, IndexEnum::BETWEEN));
  auto r = normalize_range("ACGU", -1, 1, IndexEnum::ITEM);
  CHECK(r.first == 1);
  CHECK(r.second == 3);
}
