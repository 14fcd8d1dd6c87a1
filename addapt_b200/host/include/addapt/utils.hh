// Host facade of addapt_b200 -- utilities.  Mirrors the interface of the reference's include/utils.hh
// (normalize_index / normalize_range / color / contains; reference src/utils.cc:5-56) without boost.
#pragma once
#include <algorithm>
#include <list>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace addapt {

using std::list;
using std::make_shared;
using std::map;
using std::pair;
using std::string;
using std::unordered_map;
using std::vector;

// Do indices address the items of a sequence or the gaps between them?  (reference include/utils.hh:29-32)
enum class IndexEnum { ITEM, BETWEEN };

// Python-style index: negative counts from the end; out of range throws std::string
// "no index '<i>' in '<seq>'" (reference src/utils.cc:5-27).
int normalize_index(string sequence, int index, IndexEnum meaning);
// Both ends normalised, returned in ascending order (reference src/utils.cc:29-40).
pair<int, int> normalize_range(string sequence, int start, int end, IndexEnum meaning);

enum class ColorEnum { NORMAL = 0, BLACK = 30, RED = 31, GREEN = 32, YELLOW = 33, BLUE = 34, MAGENTA = 35, CYAN = 36, WHITE = 37 };
enum class StyleEnum { NORMAL = 0, BOLD = 1, REVERSE = 2 };
string color(string str, ColorEnum color = ColorEnum::NORMAL, StyleEnum style = StyleEnum::NORMAL);

template <class Container, class Value>
bool contains(Container c, Value val) { return std::find(c.begin(), c.end(), val) != c.end(); }

// printf-free message builder used for the std::string exceptions the reference throws through boost::format
template <class... Args>
string msg(Args const&... args) {
  std::ostringstream out;
  (void)std::initializer_list<int>{(out << args, 0)...};
  return out.str();
}

}  // namespace addapt
