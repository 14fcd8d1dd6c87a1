#include "addapt/utils.hh"

namespace addapt {

int normalize_index(string sequence, int index, IndexEnum meaning) {
  const int len = (int)sequence.length();
  const int between = meaning == IndexEnum::BETWEEN ? 1 : 0;
  int resolved = index < 0 ? index + len + between : index;
  const int last = len - (between ? 0 : 1);
  if (resolved < 0 || resolved > last) throw msg("no index '", index, "' in '", sequence, "'");
  return resolved;
}

pair<int, int> normalize_range(string sequence, int start, int end, IndexEnum meaning) {
  const int a = normalize_index(sequence, start, meaning), b = normalize_index(sequence, end, meaning);
  return {std::min(a, b), std::max(a, b)};
}

string color(string str, ColorEnum color, StyleEnum style) {
  return msg("\033[", static_cast<int>(style), ";", static_cast<int>(color), "m", str, "\033[0;0m");
}

}  // namespace addapt
