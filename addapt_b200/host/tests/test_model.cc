// Device / Aptamer / Context behaviour pinned by the reference's tests/test_model.cc, restated as tables.
#include "addapt/model.hh"
#include "check.hh"

using namespace addapt;

namespace {
void expect_macrostates(Device const& d, string const& name, string const& cst) {
  int seen = 0;
  for (auto const& ms : d.macrostates()) {
    seen++;
    CHECK(ms.first == name);
    CHECK(ms.second == cst);
  }
  CHECK(seen == 1);
}
}  // namespace

TEST_CASE(model_device, "Device: sequence, python-style indexing, macrostates, copy / assign (reference tests/test_model.cc:8-41)") {
  Device four("ACGU");
  CHECK(four.seq() == "ACGU" && four.len() == 4 && four.raw_len() == 4);
  const string letters = "ACGU";
  for (int k = 0; k < 4; k++) {
    CHECK(four.seq(k) == letters[k]);            // from the front
    CHECK(four.seq(k - 4) == letters[k]);        // -4 .. -1 from the back
  }
  for (int bad : {4, -5, 17}) CHECK_THROWS(four.seq(bad));

  const vector<pair<string, string>> states = {{"a", "...."}, {"b", "(())"}};
  for (auto const& st : states) four.add_macrostate(st.first, st.second);
  CHECK_THROWS(four.macrostate("missing"));
  CHECK_THROWS(four.add_macrostate("short", "..."));     // length must match the raw sequence

  DevicePtr twin = four.copy();
  Device blank("nnnn");
  CHECK(blank.seq() == "nnnn");
  blank.assign(twin);
  for (Device const* d : {static_cast<Device const*>(&four), static_cast<Device const*>(twin.get()), static_cast<Device const*>(&blank)}) {
    CHECK(d->seq() == "ACGU");
    for (auto const& st : states) CHECK(d->macrostate(st.first) == st.second);
  }
  twin->mutate(0, 'G');                                    // copies are deep
  CHECK(four.seq() == "ACGU" && twin->seq() == "GCGU");
}

TEST_CASE(model_context, "Device embedded in a context and released again (reference tests/test_model.cc:43-89)") {
  Device d("C");
  d.add_macrostate("bp", "x");
  CHECK(d.context()->before().empty() && d.context()->after().empty());

  d.context(make_shared<Context>("A", "GU"));
  CHECK(d.len() == 4 && d.seq() == "ACGU");
  CHECK(d.seq(0) == 'A' && d.seq(1) == 'C' && d.seq(2) == 'G' && d.seq(3) == 'U');
  CHECK(d.raw_len() == 1 && d.raw_seq() == "C" && d.raw_seq(0) == 'C');      // the raw view ignores the context
  CHECK(d.macrostate("bp") == ".x..");                                      // flanks are padded with '.'
  CHECK(d.context()->before() == "A" && d.context()->after() == "GU");
  expect_macrostates(d, "bp", ".x..");

  d.remove_context();
  CHECK(d.len() == 1 && d.seq() == "C" && d.seq(0) == 'C');
  CHECK(d.raw_len() == 1 && d.raw_seq() == "C" && d.raw_seq(0) == 'C');
  CHECK(d.macrostate("bp") == "x");
  CHECK(d.context()->before().empty() && d.context()->after().empty());
  expect_macrostates(d, "bp", "x");
}

TEST_CASE(model_mutate, "Device::mutate indexes the raw sequence, either end (reference tests/test_model.cc:91-112)") {
  struct Walk { int first, step; vector<string> expect; };
  const vector<Walk> walks = {{0, 1, {"UAAA", "UUAA", "UUUA", "UUUU"}}, {-1, -1, {"AAAU", "AAUU", "AUUU", "UUUU"}}};
  for (Walk const& wk : walks) {
    Device d("AAAA");
    int at = wk.first;
    for (string const& want : wk.expect) {
      d.mutate(at, 'U');
      CHECK(d.seq() == want);
      at += wk.step;
    }
  }
  Device d("AAAA");
  for (int bad : {4, -5}) CHECK_THROWS(d.mutate(bad, 'U'));
  CHECK(d.seq() == "AAAA");
}

TEST_CASE(model_aptamer, "Aptamer accessors (reference tests/test_model.cc:114-123)") {
  const string seq = "GAUACCAGCCGAAAGGCCCUUGGCAGC", fold = "(...((.(((....)))....))...)";
  Aptamer theo(seq, fold, 0.320);
  CHECK(theo.seq() == seq && theo.fold() == fold && theo.affinity() == 0.320);
}

TEST_CASE(utils_index, "normalize_index / normalize_range (reference src/utils.cc:5-40)") {
  struct Row { int index; IndexEnum kind; int expect; };     // expect < 0: throws
  const vector<Row> rows = {{0, IndexEnum::ITEM, 0}, {3, IndexEnum::ITEM, 3}, {-1, IndexEnum::ITEM, 3}, {-4, IndexEnum::ITEM, 0},
                            {4, IndexEnum::ITEM, -1}, {-5, IndexEnum::ITEM, -1}, {4, IndexEnum::BETWEEN, 4}, {-1, IndexEnum::BETWEEN, 4},
                            {-5, IndexEnum::BETWEEN, 0}, {5, IndexEnum::BETWEEN, -1}, {-6, IndexEnum::BETWEEN, -1}};
  for (Row const& r : rows) {
    if (r.expect < 0) CHECK_THROWS(normalize_index("ACGU", r.index, r.kind));
    else CHECK(normalize_index("ACGU", r.index, r.kind) == r.expect);
  }
  auto span = normalize_range("ACGU", -1, 1, IndexEnum::ITEM);
  CHECK(span.first == 1 && span.second == 3);
  CHECK_THROWS(normalize_range("ACGU", 0, 4, IndexEnum::ITEM));
}
