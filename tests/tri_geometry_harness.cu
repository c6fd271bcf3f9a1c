// Host-side check of the lane geometry of the triangle / block kernels (chess_b200/csrc/tri_rows.cuh).  The
// geometry functions compile for host and device; this program walks every layout the launchers can choose and
// verifies, without a GPU, that the lanes of all parts of a pair tile the symmetric matrix exactly:
//   * weight(r, c) + weight(c, r) == 2 for r != c and weight(r, r) == 1  (a cell right of its band's diagonal
//     block stands for its mirror image, the diagonal block is walked as a full square, halo lanes weigh 0);
//   * the strips of a band are contiguous from lane to lane, and the lanes next to a band's first and last strip
//     are idle (the neighbour exchange of the window sums wraps onto zeros) -- lane 31 always is.
// Prints one line per (G, np) that fails and "checked N layouts" at the end; exit code 1 on any failure.
#include <cstdio>
#include <vector>
#include "../chess_b200/csrc/tri_rows.cuh"

namespace {

struct Lane { bool used; int lo, hi, cbase, cend, nv; double wl; };

int check_lanes(const Lane *ln, int np_, const char *what, int G) {
  int bad = 0;
  for (int l = 0; l < 32; ++l) {
    if (!ln[l].used) continue;
    const Lane &a = ln[l];
    const Lane &right = ln[(l + 1) & 31], &left = ln[(l + 31) & 31];
    const bool last = a.cbase + G >= a.cend;     // the matrix edge, or the end of a block's right halo lane
    if (!last && !(right.used && right.cbase == a.cbase + G && right.lo == a.lo && right.hi == a.hi)) {
      std::printf("%s G=%d np=%d lane %d: strip not continued on the next lane\n", what, G, np_, l); ++bad;
    }
    if (last && right.used) {
      std::printf("%s G=%d np=%d lane %d: lane after the last strip is in use\n", what, G, np_, l); ++bad;
    }
    const bool first = !(left.used && left.lo == a.lo && left.hi == a.hi && left.cbase == a.cbase - G);
    if (first && left.used) {
      std::printf("%s G=%d np=%d lane %d: lane before the first strip is in use\n", what, G, np_, l); ++bad;
    }
  }
  if (ln[31].used) { std::printf("%s G=%d np=%d: lane 31 in use\n", what, G, np_); ++bad; }
  return bad;
}

void add_cover(std::vector<double> &cov, int np_, const Lane &a) {
  for (int r = a.lo; r < a.hi; ++r)
    for (int g = 0; g < a.nv; ++g) cov[(size_t)r * np_ + a.cbase + g] += a.wl;
}

int check_cover(const std::vector<double> &cov, int np_, const char *what, int G) {
  for (int r = 0; r < np_; ++r) {
    if (cov[(size_t)r * np_ + r] != 1.0) {
      std::printf("%s G=%d np=%d: diagonal cell %d weighs %g\n", what, G, np_, r, cov[(size_t)r * np_ + r]);
      return 1;
    }
    for (int c = r + 1; c < np_; ++c) {
      const double w = cov[(size_t)r * np_ + c] + cov[(size_t)c * np_ + r];
      if (w != 2.0) {
        std::printf("%s G=%d np=%d: cells (%d,%d)+(%d,%d) weigh %g\n", what, G, np_, r, c, c, r, w);
        return 1;
      }
    }
  }
  return 0;
}

Lane make(const TriLane &t, int G, int np_) {
  Lane a;
  a.used = t.cbase < (1 << 20) && t.hi > t.lo;
  a.lo = t.lo; a.hi = t.hi; a.cbase = t.cbase; a.cend = t.cend; a.wl = t.wl;
  int nv = t.cend - t.cbase;
  a.nv = nv < 0 ? 0 : (nv > G ? G : nv);
  if (a.nv == 0) a.used = false;
  return a;
}

}  // namespace

int main() {
  int bad = 0, layouts = 0;
  // triangle kernels: strips of 3 .. 8 columns, compacted matrices of 1 .. 31 G (at most 248) bins, halo 3 and 5
  for (int G = 3; G <= 8; ++G) {
    for (int np_ = 1; np_ <= 31 * G && np_ <= 248; ++np_) {
      for (int halo = 3; halo <= 5; halo += 2) {
        int B1, B2, L1, L2;
        const int mode = tri_mode(np_, G, B1, B2, L1, L2);
        std::vector<double> cov((size_t)np_ * np_, 0.0);
        const int parts[2] = {mode == 1 ? 2 : 0, mode == 2 ? 1 : -1};
        for (int k = 0; k < 2; ++k) {
          if (parts[k] < 0) continue;
          Lane ln[32];
          for (int l = 0; l < 32; ++l) {
            ln[l] = make(tri_lane(parts[k], l, G, np_, B1, B2, L1, L2, halo), G, np_);
            if (ln[l].used) add_cover(cov, np_, ln[l]);
          }
          bad += check_lanes(ln, np_, parts[k] == 0 ? "tri part A" : parts[k] == 1 ? "tri part B" : "tri one pass", G);
        }
        bad += check_cover(cov, np_, "tri", G);
        if (mode == 2 && !(B1 % G == 0 && B2 % G == 0 && 0 < B1 && B1 < B2 && B2 < np_)) {
          std::printf("tri G=%d np=%d: cuts %d %d\n", G, np_, B1, B2); ++bad;
        }
        ++layouts;
      }
    }
  }
  // block kernels: strips of 5 .. 8 columns, nb x nb blocks (the launcher fixes nb from the widest pair of the
  // batch: 2 .. 7); a pair's compacted size np decides the side of its blocks, a multiple of G of at most 29 strips;
  // pairs whose blocks would be narrower than 8 bins go to the generic kernel (compare_blk.cuh)
  for (int G = 5; G <= 8; ++G) {
    for (int nb = 2; nb <= 7; ++nb) {
      for (int np_ = 8 * nb - nb + 1; np_ <= nb * 29 * G && np_ <= 1024; np_ += (np_ < 300 ? 1 : 5)) {
        const int per = (np_ + nb - 1) / nb;
        if (per < 8 || (per + G - 1) / G > 29) continue;
        const int side = ((per + G - 1) / G) * G;
        std::vector<double> cov((size_t)np_ * np_, 0.0);
        for (int bi = 0; bi < nb; ++bi) {
          for (int bj = bi; bj < nb; ++bj) {
            Lane ln[32];
            for (int l = 0; l < 32; ++l) {
              ln[l] = make(blk_lane(bi, bj, side, l, G, np_, 3), G, np_);
              if (ln[l].used) add_cover(cov, np_, ln[l]);
            }
            bad += check_lanes(ln, np_, "block", G);
          }
        }
        bad += check_cover(cov, np_, "block", G);
        ++layouts;
      }
    }
  }
  std::printf("checked %d layouts, %d failures\n", layouts, bad);
  return bad ? 1 : 0;
}
