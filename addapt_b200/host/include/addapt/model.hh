// Host facade of addapt_b200 -- the data model.  Same classes, method names and error behaviour as the
// reference's include/model.hh (Device :22-111, Aptamer :113-137, Context :139-157; src/model.cc).
// On the GPU a Device becomes a packed byte row plus static masks (addapt_b200/csrc/problem_host.h); these
// classes are what a caller of the reference API keeps using on the host.
#pragma once
#include "addapt/utils.hh"

namespace addapt {

class Device;
using DevicePtr = std::shared_ptr<Device>;
using DeviceConstPtr = std::shared_ptr<Device const>;
class Aptamer;
using AptamerConstPtr = std::shared_ptr<Aptamer const>;
class Context;
using ContextConstPtr = std::shared_ptr<Context const>;

// 5' / 3' flanks the device is embedded in (e.g. a spacer).  reference src/model.cc:121-134
class Context {
 public:
  Context(string before = "", string after = "");
  string before() const;
  string after() const;

 private:
  string my_before, my_after;
};

// Ligand-binding motif: sequence, bound fold, Kd in uM.  reference src/model.cc:102-119
class Aptamer {
 public:
  Aptamer(string seq, string fold, double affinity_uM);
  string seq() const;
  string fold() const;
  double affinity() const;

 private:
  string my_seq, my_fold;
  double my_affinity;
};

// One flat sequence (upper case = mutable, reference src/sampling.cc:170-175) with named macrostates
// (dot-bracket/x constraint strings over the RAW sequence) and an optional context.
class Device {
 public:
  Device(string seq);

  int len() const;               // with context
  string seq() const;            // before + raw + after              (src/model.cc:17-19)
  char seq(int index) const;     // python-style index, throws std::string when out of range
  int raw_len() const;
  string raw_seq() const;
  char raw_seq(int index) const;

  string macrostate(string name) const;          // padded with '.' over the context (src/model.cc:43-51); unknown name throws
  void add_macrostate(string name, string cst);  // length mismatch throws const char* (src/model.cc:59-63)
  ContextConstPtr context() const;
  void context(ContextConstPtr);
  void remove_context();
  void mutate(int raw_index, char nucleotide);   // index into the RAW sequence (src/model.cc:81-85)
  DevicePtr copy() const;
  void assign(DevicePtr other);

  // (name, context-padded constraint) pairs, in the container's order
  vector<pair<string, string>> macrostates() const;

 private:
  string my_seq;
  unordered_map<string, string> my_macrostates;
  ContextConstPtr my_context;
};

}  // namespace addapt
