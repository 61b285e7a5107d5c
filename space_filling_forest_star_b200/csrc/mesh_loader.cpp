// mesh_loader.cpp -- text meshes -> double triangle soup with the reference loader's exact semantics.
//
// Behaviour follows Obstacle<T>::ParseOBJFile / ParseMapFile / addPoint / addFacet
// (reference src/environment.h:125-223) so that both sides test identical triangles:
//   * tokens are separated by ONE space (parseString, src/primitives.h:679-695)
//   * OBJ: any line whose first token starts with 'v' ("v", "vn", "vt") contributes a vertex (:134-144);
//     an 'f' line contributes ONE triangle from its first three indices, std::stoi-style parsing so that
//     "7//3" reads as 7 (:145-154); 'o' lines never move the index offset because objId stays 0 (:155-159)
//   * .tri map: "x1 y1 x2 y2 x3 y3" per non-blank line, z = 0 (:169-195)
//   * vertex = (value + position[i]) * scale, in double (:140, :185, :198-202)
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <string>

#include "common.h"

namespace sffg {
namespace {

// first token up to the first space; `rest` receives what follows that space
void next_token(const std::string &in, std::string *tok, std::string *rest) {
  size_t pos = in.find(' ');
  if (pos == std::string::npos) {
    *tok = in;
    rest->clear();
  } else {
    *tok = in.substr(0, pos);
    *rest = in.substr(pos + 1);
  }
}

bool parse_double(const std::string &s, double *v) {
  const char *b = s.c_str();
  char *e = nullptr;
  errno = 0;
  *v = std::strtod(b, &e);
  return e != b;
}

bool parse_int(const std::string &s, long *v) {
  const char *b = s.c_str();
  char *e = nullptr;
  *v = std::strtol(b, &e, 10);
  return e != b;
}

std::string trim_ws(const std::string &s) {
  static const char *ws = " \n\r\t\f\v";
  size_t a = s.find_first_not_of(ws);
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(ws);
  return s.substr(a, b - a + 1);
}

struct Soup {
  std::vector<double> pts;   // xyz per parsed vertex
  std::vector<double> *tris;
  double bbox[6];
  double scale;

  void add_point(double c[3]) {
    for (int i = 0; i < 3; ++i) c[i] *= scale;
    pts.insert(pts.end(), c, c + 3);
    for (int i = 0; i < 3; ++i) {
      if (c[i] < bbox[2 * i]) bbox[2 * i] = c[i];
      if (c[i] > bbox[2 * i + 1]) bbox[2 * i + 1] = c[i];
    }
  }
  bool add_facet(const long idx[3]) {
    for (int i = 0; i < 3; ++i) {
      long pos = idx[i] - 1;
      if (pos < 0 || (size_t)pos * 3 + 2 >= pts.size()) return false;
      tris->insert(tris->end(), pts.begin() + pos * 3, pts.begin() + pos * 3 + 3);
    }
    return true;
  }
};

}  // namespace

int load_mesh(const char *path, int is_obj, const double position[3], double scale, std::vector<double> *tris,
              double bbox[6]) {
  std::ifstream in(path);
  if (!in.good()) return fail(SFFG_ERR_IO, std::string("cannot open mesh file ") + path);
  Soup soup;
  soup.tris = tris;
  soup.scale = scale;
  for (int i = 0; i < 3; ++i) {
    soup.bbox[2 * i] = std::numeric_limits<double>::max();
    soup.bbox[2 * i + 1] = -std::numeric_limits<double>::max();
  }
  tris->clear();
  std::string line, tok, rest;
  long lineno = 0;
  while (std::getline(in, line)) {
    ++lineno;
    if (is_obj) {
      const bool fixed = is_obj == SFFG_MESH_OBJ_FIXED;
      if (fixed) line = trim_ws(line);
      next_token(line, &tok, &rest);
      if (tok.empty()) continue;
      if (fixed ? tok == "v" : tok[0] == 'v') {
        double c[3];
        for (int i = 0; i < 3; ++i) {
          std::string cur = rest;
          next_token(cur, &tok, &rest);
          double d;
          if (!parse_double(tok, &d))
            return fail(SFFG_ERR_IO, std::string(path) + ":" + std::to_string(lineno) + ": bad vertex coordinate");
          c[i] = d + position[i];
        }
        soup.add_point(c);
      } else if (fixed && tok == "f") {
        // opt-in repair of the reference loader: every polygon is fan-triangulated (a quad gives two triangles, not one)
        // and relative (negative) indices are resolved against the vertices read so far
        std::vector<long> poly;
        while (!rest.empty()) {
          std::string cur = rest;
          next_token(cur, &tok, &rest);
          if (tok.empty()) continue;
          long v;
          if (!parse_int(tok, &v))
            return fail(SFFG_ERR_IO, std::string(path) + ":" + std::to_string(lineno) + ": bad face index");
          if (v < 0) v = (long)(soup.pts.size() / 3) + 1 + v;
          poly.push_back(v);
        }
        if (poly.size() < 3) return fail(SFFG_ERR_IO, std::string(path) + ":" + std::to_string(lineno) + ": face with fewer than 3 vertices");
        for (size_t k = 1; k + 1 < poly.size(); ++k) {
          const long idx[3] = {poly[0], poly[k], poly[k + 1]};
          if (!soup.add_facet(idx))
            return fail(SFFG_ERR_IO, std::string(path) + ":" + std::to_string(lineno) + ": face index out of range");
        }
      } else if (!fixed && tok[0] == 'f') {
        long idx[3];
        for (int i = 0; i < 3; ++i) {
          std::string cur = rest;
          next_token(cur, &tok, &rest);
          if (!parse_int(tok, &idx[i]))
            return fail(SFFG_ERR_IO, std::string(path) + ":" + std::to_string(lineno) + ": bad face index");
        }
        if (!soup.add_facet(idx))
          return fail(SFFG_ERR_IO, std::string(path) + ":" + std::to_string(lineno) + ": face index out of range");
      }
    } else {
      line = trim_ws(line);
      if (line.empty()) continue;
      long idx[3];
      for (int i = 0; i < 3; ++i) {
        double c[3] = {0.0, 0.0, 0.0};
        for (int j = 0; j < 2; ++j) {
          std::string cur = line;
          next_token(cur, &tok, &line);
          double d;
          if (!parse_double(tok, &d))
            return fail(SFFG_ERR_IO, std::string(path) + ":" + std::to_string(lineno) + ": bad map coordinate");
          c[j] = d + position[j];
        }
        soup.add_point(c);
        idx[i] = (long)(soup.pts.size() / 3);
      }
      soup.add_facet(idx);
    }
  }
  if (bbox)
    for (int i = 0; i < 6; ++i) bbox[i] = soup.bbox[i];
  return SFFG_OK;
}

}  // namespace sffg
