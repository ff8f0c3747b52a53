// libint_b200_basis.hpp -- header-only C++ mirror of libint2::Atom / read_dotxyz / BasisSet for the hosts of the
// B200 library (plain C++17; needs only lb200_shell_renorm from -llibint_b200, no GPU).
//
// What it mirrors (evaleev/libint, file:line):
//   libint2::Atom, constants            include/libint2/atom.h:39-67
//   libint2::read_dotxyz                include/libint2/atom.h:83-160,208-220 (XYZ file in Angstrom -> bohr)
//   libint2::BasisSet                   include/libint2/basis.h.in:89-617: std::vector<Shell> with
//                                       BasisSet(name, atoms, throw_if_no_match) (:134-180), set_pure (:257),
//                                       nbf / max_nprim / max_l / shell2bf (:265-281), shell2atom / atom2shell
//                                       (:284-333), the Gaussian Cartesian-d convention (:368-386), the
//                                       aug-cc-pVXZ decomposition (:388-400), data_path() (:412-450)
// Differences, all on the data side: the basis library is the packed JSON re-encoding of the reference's
// lib/basis/*.g94 files (libint_b200/data/basis/<name>.json, '*' spelled 's' in file names; tools/pack_basis.py),
// found through LIBINT_B200_DATA_PATH (the directory that holds basis/), else the LIBINT_B200_DATADIR macro.
// Shells with several contractions in the .g94 file (SP shells) are split by the packer, as the reference's reader
// does (basis.h.in:520-560).
#ifndef LIBINT_B200_BASIS_HPP
#define LIBINT_B200_BASIS_HPP

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <istream>
#include <map>
#include <sstream>
#include <system_error>

#include "libint_b200.hpp"

namespace libint_b200 {

namespace constants {   // atom.h:46-67
struct codata_2018 { static constexpr double bohr_to_angstrom = 0.529177210903; };
struct codata_2010 { static constexpr double bohr_to_angstrom = 0.52917721092; };
}  // namespace constants

/// libint2::Atom (atom.h:39-42): coordinates in bohr
struct Atom {
  int atomic_number;
  double x, y, z;
};

namespace detail {
inline std::string lower(std::string s) {
  for (auto& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}
inline int element_to_Z(const std::string& symbol) {
  static const char* sym[] = {"",   "h",  "he", "li", "be", "b",  "c",  "n",  "o",  "f",  "ne", "na", "mg",
                              "al", "si", "p",  "s",  "cl", "ar", "k",  "ca", "sc", "ti", "v",  "cr", "mn",
                              "fe", "co", "ni", "cu", "zn", "ga", "ge", "as", "se", "br", "kr"};
  const std::string s = lower(symbol);
  for (int z = 1; z <= 36; ++z)
    if (s == sym[z]) return z;
  return -1;
}

// reader of the packed basis files: {"name": ..., "shells": {"Z": [[l, [exps], [coefs]], ...]}}
struct json_cursor {
  const std::string& s;
  size_t i = 0;
  explicit json_cursor(const std::string& text) : s(text) {}
  void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
  bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
  void expect(char c) { if (!eat(c)) throw std::runtime_error(std::string("basis file: expected '") + c + "'"); }
  std::string str() {
    expect('"');
    std::string r;
    while (i < s.size() && s[i] != '"') r += s[i++];
    ++i;
    return r;
  }
  double num() {
    ws();
    char* end = nullptr;
    const double v = std::strtod(s.c_str() + i, &end);
    if (end == s.c_str() + i) throw std::runtime_error("basis file: number expected");
    i = (size_t)(end - s.c_str());
    return v;
  }
  std::vector<double> numlist() {
    std::vector<double> v;
    expect('[');
    if (eat(']')) return v;
    do v.push_back(num()); while (eat(','));
    expect(']');
    return v;
  }
};
struct raw_shell { int l; std::vector<double> exps, coefs; };
using element_library = std::map<int, std::vector<raw_shell>>;

inline element_library read_packed_basis(const std::string& path, std::string* name = nullptr) {
  std::ifstream is(path);
  if (!is) throw std::ios_base::failure("cannot open basis file " + path);
  std::stringstream ss;
  ss << is.rdbuf();
  const std::string text = ss.str();
  json_cursor j(text);
  element_library out;
  j.expect('{');
  do {
    const std::string key = j.str();
    j.expect(':');
    if (key == "name") {
      const std::string n = j.str();
      if (name) *name = n;
    } else if (key == "shells") {
      j.expect('{');
      do {
        const int Z = std::atoi(j.str().c_str());
        j.expect(':');
        j.expect('[');
        do {
          raw_shell sh;
          j.expect('[');
          sh.l = (int)j.num();
          j.expect(',');
          sh.exps = j.numlist();
          j.expect(',');
          sh.coefs = j.numlist();
          j.expect(']');
          out[Z].push_back(std::move(sh));
        } while (j.eat(','));
        j.expect(']');
      } while (j.eat(','));
      j.expect('}');
    } else {
      throw std::runtime_error("basis file: unknown key " + key);
    }
  } while (j.eat(','));
  return out;
}
}  // namespace detail

/// libint2::read_dotxyz (atom.h:208-220): "natoms \n comment \n symbol x y z ..." in Angstrom -> atoms in bohr
inline std::vector<Atom> read_dotxyz(std::istream& is,
                                     double bohr_to_angstrom = constants::codata_2018::bohr_to_angstrom) {
  std::string line;
  std::getline(is, line);
  const long natom = std::atol(line.c_str());
  if (!is || natom < 0) throw std::logic_error("read_dotxyz: expected the number of atoms on the first line");
  std::getline(is, line);   // comment
  std::vector<Atom> atoms;
  const double angstrom_to_bohr = 1 / bohr_to_angstrom;
  for (long a = 0; a < natom; ++a) {
    std::getline(is, line);
    std::istringstream ls(line);
    std::string el;
    Atom at{};
    ls >> el >> at.x >> at.y >> at.z;
    at.atomic_number = detail::element_to_Z(el);
    if (!ls || at.atomic_number < 0)
      throw std::logic_error("read_dotxyz: bad atom line / element \"" + el + "\"");   // atom.h:125-135
    at.x *= angstrom_to_bohr;   // atom.h:123-130 multiplies by the reciprocal
    at.y *= angstrom_to_bohr;
    at.z *= angstrom_to_bohr;
    atoms.push_back(at);
  }
  return atoms;
}

/// libint2::BasisSet (basis.h.in:89): the shells of a named basis placed on the atoms of a molecule
class BasisSet : public std::vector<Shell> {
 public:
  using base_type = std::vector<Shell>;

  BasisSet() = default;
  /// from shells (basis.h.in:108-109)
  BasisSet(base_type shells) : base_type(std::move(shells)) { init(); }   // NOLINT: implicit like the reference's
  /// BasisSet(name, atoms) (basis.h.in:134-180): every atom receives its element's shells of every component of
  /// the named basis, in file order, moved to the atom
  BasisSet(std::string name, const std::vector<Atom>& atoms, bool throw_if_no_match = false) : name_(std::move(name)) {
    const std::string canonical = canonicalize_name(name_);
    std::vector<std::string> files;
    for (const std::string& comp : decompose_name_into_components(canonical))
      files.push_back(data_path() + "/" + file_stem(comp) + ".json");
    build(files, gaussian_cartesian_d_convention(canonical), atoms, throw_if_no_match);
  }
  /// the same from explicit component files (what the command-line driver is handed); the Cartesian-d rule goes
  /// by the name recorded in the first file
  static BasisSet from_files(const std::vector<std::string>& files, const std::vector<Atom>& atoms,
                             bool throw_if_no_match = false) {
    if (files.empty()) throw std::logic_error("BasisSet::from_files: no basis file given");
    BasisSet bs;
    detail::read_packed_basis(files[0], &bs.name_);
    bs.build(files, gaussian_cartesian_d_convention(canonicalize_name(bs.name_)), atoms, throw_if_no_match);
    return bs;
  }

  const std::string& name() const { return name_; }
  /// forces solid harmonics / Cartesian Gaussians: sets the flag of every shell, s and p included, exactly as
  /// basis.h.in:257-262 does (the library honours a pure p shell, tests/test_gpu_fock.py)
  void set_pure(bool solid) {
    for (Shell& s : *this) s.pure = solid;
    init();
  }
  long nbf() const { return nbf_; }
  size_t max_nprim() const { return max_nprim_; }
  long max_l() const { return max_l_; }
  const std::vector<size_t>& shell2bf() const { return shell2bf_; }
  std::vector<long> shell2atom(const std::vector<Atom>& atoms) const { return shell2atom(*this, atoms, false); }
  std::vector<std::vector<long>> atom2shell(const std::vector<Atom>& atoms) const { return atom2shell(atoms, *this); }

  /// shell -> the atom whose position equals the shell's origin bit for bit, -1 if none (basis.h.in:300-313)
  static std::vector<long> shell2atom(const std::vector<Shell>& shells, const std::vector<Atom>& atoms,
                                      bool throw_if_no_match = false) {
    std::vector<long> result;
    result.reserve(shells.size());
    for (const Shell& s : shells) {
      long hit = -1;
      for (size_t a = 0; a < atoms.size() && hit < 0; ++a)
        if (s.O[0] == atoms[a].x && s.O[1] == atoms[a].y && s.O[2] == atoms[a].z) hit = (long)a;
      if (hit < 0 && throw_if_no_match) throw std::logic_error("shell2atom: no matching atom found");
      result.push_back(hit);
    }
    return result;
  }
  /// atom -> the shells centred on it (basis.h.in:315-333)
  static std::vector<std::vector<long>> atom2shell(const std::vector<Atom>& atoms, const std::vector<Shell>& shells) {
    std::vector<std::vector<long>> result(atoms.size());
    for (size_t a = 0; a < atoms.size(); ++a)
      for (size_t s = 0; s < shells.size(); ++s)
        if (shells[s].O[0] == atoms[a].x && shells[s].O[1] == atoms[a].y && shells[s].O[2] == atoms[a].z)
          result[a].push_back((long)s);
    return result;
  }

  /// the directory holding the packed basis files (basis.h.in:412-450): $LIBINT_B200_DATA_PATH/basis, else
  /// LIBINT_B200_DATADIR "/basis"
  static std::string data_path() {
    std::string path;
    if (const char* env = std::getenv("LIBINT_B200_DATA_PATH")) {
      path = env;
    } else {
#ifdef LIBINT_B200_DATADIR
      path = LIBINT_B200_DATADIR;
#else
      throw std::system_error(std::make_error_code(std::errc::no_such_file_or_directory),
                              "BasisSet::data_path: set LIBINT_B200_DATA_PATH (the directory that holds basis/)");
#endif
    }
    return path + "/basis";
  }
  /// lower case (basis.h.in:347-366)
  static std::string canonicalize_name(const std::string& name) { return detail::lower(name); }
  /// basis.h.in:368-386: the 3-21G / 4-31G / 6-31G families use Cartesian d shells (Gaussian's convention)
  static bool gaussian_cartesian_d_convention(const std::string& n) {
    if (n.rfind("3-21", 0) == 0 || n.rfind("4-31g", 0) == 0) return true;
    if (n.rfind("6-31", 0) == 0 && n.size() > 4 && n[4] != '1') {
      const size_t g = n.find('g');
      if (g == std::string::npos) return false;
      if (g + 1 == n.size()) return true;
      if (n[g + 1] == '*' || n[g + 1] == 's') return true;
    }
    return false;
  }
  /// basis.h.in:388-400: aug-cc-pVXZ* = cc-pVXZ* + augmentation-cc-pVXZ*, except the -cabs sets
  static std::vector<std::string> decompose_name_into_components(const std::string& name) {
    if (name.rfind("aug-cc-pv", 0) == 0 && name.find("cabs") == std::string::npos)
      return {name.substr(4), "augmentation-" + name.substr(4)};
    return {name};
  }

 private:
  std::string name_;
  long nbf_ = -1;
  size_t max_nprim_ = 0;
  long max_l_ = -1;
  std::vector<size_t> shell2bf_;

  static std::string file_stem(std::string n) {   // '*' is spelled 's' in the packed file names
    for (auto& c : n)
      if (c == '*') c = 's';
    return n;
  }
  void build(const std::vector<std::string>& files, bool cartesian_d, const std::vector<Atom>& atoms,
             bool throw_if_no_match) {
    std::vector<detail::element_library> comps;
    for (const std::string& f : files) comps.push_back(detail::read_packed_basis(f));
    for (const Atom& at : atoms)
      for (size_t c = 0; c < comps.size(); ++c) {
        const auto it = comps[c].find(at.atomic_number);
        if (it == comps[c].end() || it->second.empty()) {
          if (throw_if_no_match)
            throw std::logic_error("did not find the basis for Z = " + std::to_string(at.atomic_number) + " in " + files[c]);
          continue;
        }
        for (const detail::raw_shell& r : it->second)   // pure iff l > 1, or l > 2 under the Cartesian-d convention
          emplace_back(r.exps, r.l, cartesian_d ? r.l > 2 : r.l > 1, r.coefs, std::array<double, 3>{{at.x, at.y, at.z}});
      }
    init();
  }
  void init() {   // basis.h.in:340-345
    nbf_ = 0;
    max_nprim_ = 0;
    max_l_ = -1;
    shell2bf_.clear();
    for (const Shell& s : *this) {
      shell2bf_.push_back((size_t)nbf_);
      nbf_ += (long)s.size();
      max_nprim_ = std::max(max_nprim_, s.nprim());
      max_l_ = std::max(max_l_, (long)s.l);
    }
  }
};

}  // namespace libint_b200

#endif  // LIBINT_B200_BASIS_HPP
