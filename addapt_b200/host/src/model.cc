#include "addapt/model.hh"

namespace addapt {

Context::Context(string before, string after) : my_before(before), my_after(after) {}
string Context::before() const { return my_before; }
string Context::after() const { return my_after; }

Aptamer::Aptamer(string seq, string fold, double affinity_uM) : my_seq(seq), my_fold(fold), my_affinity(affinity_uM) {}
string Aptamer::seq() const { return my_seq; }
string Aptamer::fold() const { return my_fold; }
double Aptamer::affinity() const { return my_affinity; }

Device::Device(string seq) : my_seq(seq), my_context(make_shared<Context>()) {}

string Device::seq() const { return my_context->before() + my_seq + my_context->after(); }
int Device::len() const { return (int)(my_context->before().size() + my_seq.size() + my_context->after().size()); }
char Device::seq(int index) const {
  const string full = seq();
  return full[normalize_index(full, index, IndexEnum::ITEM)];
}
int Device::raw_len() const { return (int)my_seq.size(); }
string Device::raw_seq() const { return my_seq; }
char Device::raw_seq(int index) const { return my_seq[normalize_index(my_seq, index, IndexEnum::ITEM)]; }

string Device::macrostate(string name) const {
  auto it = my_macrostates.find(name);
  if (it == my_macrostates.end()) throw msg("no macrostate '", name, "'");
  return string(my_context->before().size(), '.') + it->second + string(my_context->after().size(), '.');
}

void Device::add_macrostate(string name, string cst) {
  if (cst.size() != my_seq.size()) throw "constraint length doesn't match sequence length";
  my_macrostates[name] = cst;
}

ContextConstPtr Device::context() const { return my_context; }
void Device::context(ContextConstPtr ctx) { my_context = ctx; }
void Device::remove_context() { my_context = make_shared<Context>(); }

void Device::mutate(int raw_index, char nucleotide) { my_seq[normalize_index(my_seq, raw_index, IndexEnum::ITEM)] = nucleotide; }

DevicePtr Device::copy() const {
  DevicePtr twin = make_shared<Device>(my_seq);
  twin->my_macrostates = my_macrostates;
  twin->my_context = my_context;
  return twin;
}

void Device::assign(DevicePtr other) {
  my_seq = other->my_seq;
  my_macrostates = other->my_macrostates;
  my_context = other->my_context;
}

vector<pair<string, string>> Device::macrostates() const {
  vector<pair<string, string>> out;
  for (auto const& kv : my_macrostates) out.emplace_back(kv.first, macrostate(kv.first));
  return out;
}

}  // namespace addapt
