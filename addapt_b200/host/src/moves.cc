#include <cctype>

#include "addapt/moves.hh"

namespace addapt {

bool can_be_mutated(DeviceConstPtr device, int position) { return std::isupper((unsigned char)device->seq()[position]) != 0; }

bool can_be_freely_mutated(DeviceConstPtr device, int position) {
  if (!can_be_mutated(device, position)) return false;
  for (auto const& ms : device->macrostates())
    if (ms.second[position] == ')') return false;          // mutated together with its 5' partner instead
  return true;
}

void mutate_recursively(DevicePtr device, int const position, char const mutation) {
  vector<bool> done(device->len(), false);
  mutate_recursively(device, position, mutation, done);
}

void mutate_recursively(DevicePtr device, int const position, char const mutation, vector<bool>& already_mutated) {
  device->mutate(position, mutation);
  already_mutated[position] = true;

  for (auto const& ms : device->macrostates()) {
    string const& brackets = ms.second;
    const char here = brackets[position];
    if (here != '(' && here != ')') continue;
    // walk towards the partner, counting nesting depth
    const int dir = here == '(' ? 1 : -1;
    const char same = here, other = here == '(' ? ')' : '(';
    int depth = 1, partner = position;
    while (depth != 0) {
      partner += dir;
      if (partner < 0 || partner >= (int)brackets.size())
        throw msg("mismatched base-pair in '", ms.first, "' macrostate: '", brackets, "'");
      if (brackets[partner] == same) depth++;
      else if (brackets[partner] == other) depth--;
    }
    if (!can_be_mutated(device, partner))
      throw msg("position '", position, "' can be mutated, but it's base-paired to position '", partner, "' which can't be.");
    // Watson-Crick only: GU pairs would bias the nucleotide distribution (reference src/sampling.cc:220-222)
    const char complement = COMPLEMENTARY_NUCS.at(mutation);
    if (!already_mutated[partner]) mutate_recursively(device, partner, complement, already_mutated);
    else if (device->seq()[partner] != complement) throw "no way to satisfy all base pairing constraints.";
  }
}

UnbiasedMutationMove::UnbiasedMutationMove() {}

void UnbiasedMutationMove::apply(DevicePtr device, std::mt19937& rng) const {
  vector<int> free_positions;
  for (int i = 0; i < device->len(); i++)
    if (can_be_freely_mutated(device, i)) free_positions.push_back(i);
  const int where = free_positions[std::uniform_int_distribution<>(0, (int)free_positions.size() - 1)(rng)];
  const char what = "ACGU"[std::uniform_int_distribution<>(0, 3)(rng)];
  mutate_recursively(device, where, what);
}

}  // namespace addapt
