#include "addapt/config.hh"

#include <fstream>
#include <regex>

#include "addapt/b200.hh"

namespace addapt {

namespace {

// One parsed file: top-level key -> scalar, or key -> ordered (sub-key, value) list.
struct YamlDoc {
  map<string, string> scalars;
  map<string, vector<pair<string, string>>> maps;
};

string strip(string s) {
  // drop comments that start at a word boundary and surrounding blanks / quotes
  bool quoted = false;
  for (size_t k = 0; k < s.size(); k++) {
    if (s[k] == '"' || s[k] == '\'') quoted = !quoted;
    if (s[k] == '#' && !quoted && (k == 0 || s[k - 1] == ' ' || s[k - 1] == '\t')) { s = s.substr(0, k); break; }
  }
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  if (a == string::npos) return "";
  s = s.substr(a, b - a + 1);
  if (s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front()) s = s.substr(1, s.size() - 2);
  return s;
}

YamlDoc parse_yaml(string const& path) {
  std::ifstream in(path);
  if (!in.is_open()) throw msg("couldn't open '", path, "'");
  YamlDoc doc;
  string line, section;
  while (std::getline(in, line)) {
    if (strip(line).empty()) continue;
    const bool nested = line[0] == ' ' || line[0] == '\t';
    string body = strip(line);
    size_t colon = body.find(':');
    if (colon == string::npos) throw msg("can't parse '", line, "' in '", path, "'");
    string key = strip(body.substr(0, colon)), value = strip(body.substr(colon + 1));
    if (!nested) {
      if (doc.scalars.count(key) || doc.maps.count(key)) throw msg("section '", key, "' defined twice in '", path, "'");
      if (value.empty()) { section = key; doc.maps[key]; }
      else if (value.front() == '{' && value.back() == '}') {       // inline map {a: x, b: y}
        section = "";
        auto& items = doc.maps[key];
        string inner = value.substr(1, value.size() - 2);
        size_t pos = 0;
        int depth = 0;
        string item;
        auto flush = [&]() {
          string it = strip(item);
          item.clear();
          if (it.empty()) return;
          size_t c2 = it.find(':');
          if (c2 == string::npos) throw msg("can't parse '", it, "' in '", path, "'");
          items.emplace_back(strip(it.substr(0, c2)), strip(it.substr(c2 + 1)));
        };
        for (; pos < inner.size(); pos++) {
          char ch = inner[pos];
          if (ch == '[') depth++;
          if (ch == ']') depth--;
          if (ch == ',' && depth == 0) flush(); else item += ch;
        }
        flush();
      } else { section = ""; doc.scalars[key] = value; }
    } else {
      if (section.empty()) throw msg("unexpected indentation at '", line, "' in '", path, "'");
      doc.maps[section].emplace_back(key, value);
    }
  }
  return doc;
}

enum Requirement { REQUIRED, OPTIONAL };

// The section may live in any ONE of the files; twice is an error (reference src/config.cc:24-50).
struct Section { bool found = false; bool is_map = false; string scalar; vector<pair<string, string>> items; };

Section find_section(vector<string> const& files, string name, Requirement req) {
  Section hit;
  string where;
  for (string const& path : files) {
    YamlDoc doc = parse_yaml(path);
    const bool s = doc.scalars.count(name) > 0, m = doc.maps.count(name) > 0;
    if (!s && !m) continue;
    if (hit.found) throw msg("section '", name, "' found in both '", where, "' and '", path, "'");
    hit.found = true;
    hit.is_map = m;
    if (s) hit.scalar = doc.scalars[name]; else hit.items = doc.maps[name];
    where = path;
  }
  if (!hit.found && req == REQUIRED) throw msg("no '", name, "' section found in any config file");
  return hit;
}

string item_of(Section const& sec, string key, string section_name) {
  for (auto const& kv : sec.items)
    if (kv.first == key) return kv.second;
  throw msg("no '", key, "' in section '", section_name, "'");
}

pair<string, string> two_list(string value) {
  value = strip(value);
  if (value.size() < 2 || value.front() != '[' || value.back() != ']') throw msg("expected a [5', 3'] list, got '", value, "'");
  string inner = value.substr(1, value.size() - 2);
  size_t comma = inner.find(',');
  if (comma == string::npos) throw msg("expected two entries in '", value, "'");
  return {strip(inner.substr(0, comma)), strip(inner.substr(comma + 1))};
}

}  // namespace

DevicePtr device_from_yaml(vector<string> config_files) {
  Section seq = find_section(config_files, "sequence", REQUIRED);
  Section macro = find_section(config_files, "macrostates", REQUIRED);
  DevicePtr device = make_shared<Device>(seq.scalar);
  for (auto const& kv : macro.items) device->add_macrostate(kv.first, kv.second);
  return device;
}

ScoreFunctionPtr scorefxn_from_yaml(vector<string> config_files) {
  auto scorefxn = make_shared<B200ScoreFunction>();
  Section objective = find_section(config_files, "objective", REQUIRED);
  scorefxn->add_term(score_term_from_str(ConditionEnum::APO, item_of(objective, "apo", "objective")));
  scorefxn->add_term(score_term_from_str(ConditionEnum::HOLO, item_of(objective, "holo", "objective")));
  Section apt = find_section(config_files, "aptamer", REQUIRED);
  scorefxn->aptamer(make_shared<Aptamer>(item_of(apt, "sequence", "aptamer"), item_of(apt, "fold", "aptamer"),
                                         std::stod(item_of(apt, "affinity", "aptamer"))));
  // optional, beyond the reference's schema: terms over base-pair probabilities, `pair_objective: {name: "<apo|holo> [not] pair <i> <j>
  // [x <weight>]", name2: "<apo|holo> [not] density <first>-<last> [x <weight>]"}` with 0-based positions of the raw sequence
  Section bp = find_section(config_files, "pair_objective", OPTIONAL);
  for (auto const& kv : bp.items) scorefxn->add_term(bp_term_from_str(kv.second));
  Section ctx = find_section(config_files, "contexts", OPTIONAL);
  for (auto const& kv : ctx.items) {
    auto flanks = two_list(kv.second);
    scorefxn->add_context(kv.first, make_shared<Context>(flanks.first, flanks.second));
  }
  return scorefxn;
}

ScoreTermPtr score_term_from_str(ConditionEnum condition, string spec) {
  static const std::regex pattern("(not )?(\\w+)");
  std::smatch match;
  if (std::regex_match(spec, match, pattern))
    return make_shared<MacrostateProbTerm>(match[2], condition, match[1].matched ? FavorableEnum::NO : FavorableEnum::YES);
  throw msg("can't understand objective: '", spec, "'");
}

ScoreTermPtr bp_term_from_str(string spec) {
  static const std::regex pattern("(apo|holo) (not )?(pair|density) ([0-9]+)[ -]([0-9]+)(?: x ([0-9.e+-]+))?");
  std::smatch m;
  if (!std::regex_match(spec, m, pattern)) throw msg("can't understand pair objective: '", spec, "'");
  const ConditionEnum cond = m[1] == "apo" ? ConditionEnum::APO : ConditionEnum::HOLO;
  const FavorableEnum fav = m[2].matched ? FavorableEnum::NO : FavorableEnum::YES;
  const int a = std::stoi(m[4]), b = std::stoi(m[5]);
  ScoreTermPtr term;
  if (m[3] == "pair") term = make_shared<PairProbTerm>(a, b, cond, fav);
  else {
    if (b < a) throw msg("empty position range in '", spec, "'");
    vector<int> pos;
    for (int k = a; k <= b; k++) pos.push_back(k);
    term = make_shared<PairedDensityTerm>(pos, cond, fav);
  }
  if (m[6].matched) term->weight(std::stod(m[6]));
  return term;
}

ThermostatPtr thermostat_from_yaml(vector<string> config_files) {
  Section sec = find_section(config_files, "thermostat", OPTIONAL);
  return thermostat_from_str(sec.found ? sec.scalar : "1");
}

ThermostatPtr thermostat_from_str(string spec) {
  static const string number = "([0-9.e+-]+)";
  static const std::regex fixed(number);
  static const std::regex annealing(number + " to " + number + " in ([0-9]+) steps");
  static const std::regex autoscale("auto(?:\\s+([0-9.]+)%(?:\\s+([0-9]+)(?:\\s+" + number + ")?)?)?");
  std::smatch m;
  if (std::regex_match(spec, m, fixed)) return make_shared<FixedThermostat>(std::stod(m[1]));
  if (std::regex_match(spec, m, annealing)) return make_shared<AnnealingThermostat>(std::stoi(m[3]), std::stod(m[1]), std::stod(m[2]));
  if (std::regex_match(spec, m, autoscale)) {
    const double rate = std::stod(m[1].length() ? m[1].str() : string("50")) / 100;
    const unsigned period = (unsigned)std::stod(m[2].length() ? m[2].str() : string("100"));
    const double t0 = std::stod(m[3].length() ? m[3].str() : string("1"));
    return make_shared<AutoScalingThermostat>(rate, period, t0);
  }
  throw msg("can't make a thermostat from '", spec, "'");
}

}  // namespace addapt
