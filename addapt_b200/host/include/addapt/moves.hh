// Host facade of addapt_b200 -- mutation rules and moves (reference include/sampling.hh:107-143, src/sampling.cc:170-303).
#pragma once
#include "addapt/mc_step.hh"

namespace addapt {

map<char, char> const COMPLEMENTARY_NUCS = {{'A', 'U'}, {'G', 'C'}, {'C', 'G'}, {'U', 'A'}};

// upper case = mutable (reference src/sampling.cc:170-175)
bool can_be_mutated(DeviceConstPtr, int position);
// mutable and not the 3' end ')' of a pair in any macrostate (reference src/sampling.cc:177-193)
bool can_be_freely_mutated(DeviceConstPtr, int position);
// point mutation plus Watson-Crick complements through every macrostate's bracket partners
// (reference src/sampling.cc:205-282).  Throws std::string for an unbalanced bracket or an immutable partner,
// const char* when two constraints contradict each other.
void mutate_recursively(DevicePtr, int const position, char const mutation);
void mutate_recursively(DevicePtr, int const position, char const mutation, vector<bool>& already_mutated);

class Move {
 public:
  virtual ~Move() {}
  virtual string name() const = 0;
  virtual void apply(DevicePtr, std::mt19937&) const = 0;
};

// uniform freely-mutable position x uniform ACGU, identity included (reference src/sampling.cc:287-303)
class UnbiasedMutationMove : public Move {
 public:
  UnbiasedMutationMove();
  string name() const { return "UnbiasedMutation"; }
  void apply(DevicePtr, std::mt19937&) const;
};

}  // namespace addapt
