/*
 * oracle_aos.cpp -- TEST INFRASTRUCTURE ONLY (see picnic_oracle.h).
 *
 * The same arithmetic as orc_advance_particles_iteratively + orc_deposit_current, driven the way the reference drives it
 * in memory ("faithful" CPU number): particles are 176-byte objects (JustinsParticle, JustinsParticle.H:198-213)
 * in a doubly linked list (Chombo List<P>); every particle-Picard pass walks the list of unconverged particles, calls the
 * per-particle gather kernel once per particle (the reference: one Fortran call per particle and field set,
 * MeshInterpI.H:537-690), stores E_p / B_p back into the particle, calls the Boris kernel on the stored fields
 * (PicSpeciesUtils::applyForces), and stepNormTransfer moves the particle between the lists (PicChargedSpecies.cpp:658-733,
 * 1614-1716); setCurrentDensity then walks the list again with one deposit call per particle (:3184-3253).
 * Results per particle are bit-identical to the SoA oracle (tests/test_oracle_invariants.py); only J's summation order
 * follows the list order.  Non-relativistic build only.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "picnic_oracle.h"

namespace {

struct AosParticle {      /* field for field what JustinsParticle holds in 2D: 176 bytes */
  void *vptr;
  double position[2];
  uint64_t id;
  int kill_tag, num_suborbits;
  double weight;
  double velocity[3];
  double electric_field[3];
  double magnetic_field[3];
  double pos_virt[2];
  double position_old[2];
  double velocity_old[3];
};
static_assert(sizeof(AosParticle) == 176, "JustinsParticle is 176 bytes in 2D");

struct Node {
  Node *prev, *next;
  AosParticle p;
};
struct List {
  Node *head = nullptr, *tail = nullptr;
  long len = 0;
  void append(Node *n) {
    n->prev = tail;
    n->next = nullptr;
    if (tail) tail->next = n;
    else head = n;
    tail = n;
    ++len;
  }
  void remove(Node *n) {
    if (n->prev) n->prev->next = n->next;
    else head = n->next;
    if (n->next) n->next->prev = n->prev;
    else tail = n->prev;
    --len;
  }
};
struct Aos {
  int D;
  long n;
  std::vector<Node *> nodes;   /* ownership, creation (= id) order */
  List main;
};

/* stepNormTransfer for one particle (PicChargedSpecies.cpp:658-733) */
bool step_norm(const orc_geom *g, AosParticle &q, double cnormDt, double rtol, bool reverse) {
  const double cnormHalfDt = 0.5 * cnormDt;
  double dxp[2] = {0.0, 0.0};
  double rel_diff_max = 0.0;
  for (int d = 0; d < g->D; ++d) {
    const double dxp0 = q.position[d] - q.position_old[d];
    dxp[d] = q.velocity[d] * cnormHalfDt;
    const double rel_diff_dir = std::fabs(dxp0 - dxp[d]) / g->dx[d];
    rel_diff_max = std::max(rel_diff_max, rel_diff_dir);
  }
  if (reverse) {
    if (rel_diff_max < rtol) return true;
    for (int d = 0; d < g->D; ++d) q.position[d] = q.position_old[d] + dxp[d];
    return false;
  }
  for (int d = 0; d < g->D; ++d) q.position[d] = q.position_old[d] + dxp[d];
  return !(rel_diff_max >= rtol);
}

}  // namespace

/* scattered != 0: the nodes of consecutive particles lie at pseudo-random places of the heap, as after many steps of
 * list transfers, migration and re-binning in the reference (the list order stays the particle order) */
extern "C" void *orc_aos_create_ex(int D, long n, const double *x, const double *xold, const double *v,
                                   const double *vold, const double *w, int scattered) {
  if (orc_get_relativistic()) return nullptr;
  Aos *h = new Aos();
  h->D = D;
  h->n = n;
  h->nodes.reserve(n);
  std::vector<Node *> slot(n);
  for (long p = 0; p < n; ++p) slot[p] = new Node();
  if (scattered) {   /* Fisher-Yates with a fixed 64-bit LCG */
    uint64_t st = 0x9E3779B97F4A7C15ull;
    for (long p = n - 1; p > 0; --p) {
      st = st * 6364136223846793005ull + 1442695040888963407ull;
      const long q = (long)((st >> 33) % (uint64_t)(p + 1));
      std::swap(slot[p], slot[q]);
    }
  }
  for (long p = 0; p < n; ++p) {
    Node *nd = slot[p];
    AosParticle &q = nd->p;
    q.vptr = nullptr;
    q.id = (uint64_t)p;
    q.kill_tag = q.num_suborbits = 0;
    q.weight = w[p];
    for (int d = 0; d < 2; ++d) {
      q.position[d] = d < D ? x[d * n + p] : 0.0;
      q.position_old[d] = d < D ? xold[d * n + p] : 0.0;
      q.pos_virt[d] = 0.0;
    }
    for (int c = 0; c < 3; ++c) {
      q.velocity[c] = v[c * n + p];
      q.velocity_old[c] = vold[c * n + p];
      q.electric_field[c] = q.magnetic_field[c] = 0.0;
    }
    h->nodes.push_back(nd);
    h->main.append(nd);
  }
  return h;
}

extern "C" void *orc_aos_create(int D, long n, const double *x, const double *xold, const double *v,
                                const double *vold, const double *w) {
  return orc_aos_create_ex(D, n, x, xold, v, vold, w, 0);
}

extern "C" void orc_aos_destroy(void *hv) {
  Aos *h = static_cast<Aos *>(hv);
  if (!h) return;
  for (Node *n : h->nodes) delete n;
  delete h;
}

/* advanceParticlesIteratively (PicChargedSpecies.cpp:1614-1716) + setCurrentDensity (:3184-3253) over the lists.
 * J accumulates (no charge/volume_scale factor, like orc_deposit_current). */
extern "C" int orc_aos_advance_deposit(void *hv, const orc_geom *g, int interp, const orc_fab *E, const orc_fab *B,
                                       double fnorm, double cnormDt, double rtol, int iter_max, orc_fab *J,
                                       long *num_apply_its, long *num_unconverged) {
  Aos *h = static_cast<Aos *>(hv);
  int rc = 0;
  long apply_its = 0;
  auto gather_list = [&](List &L) {   /* interpolateFieldsToParticles: one kernel call per particle */
    for (Node *n = L.head; n; n = n->next)
      if (orc_gather(g, interp, 1, n->p.position, n->p.position_old, E, B, n->p.electric_field, n->p.magnetic_field))
        rc = -1;
  };
  auto boris_list = [&](List &L) {    /* applyForces on the stored fields */
    for (Node *n = L.head; n; n = n->next) {
      double u[3];
      orc_boris(1, u, n->p.velocity_old, n->p.electric_field, n->p.magnetic_field, fnorm, cnormDt, 1);
      n->p.velocity[0] = u[0];
      n->p.velocity[1] = u[1];
      n->p.velocity[2] = u[2];
    }
  };
  List temp;
  gather_list(h->main);
  boris_list(h->main);
  apply_its += h->main.len;
  for (Node *n = h->main.head; n;) {   /* stepNormTransfer(main -> temp) */
    Node *next = n->next;
    if (!step_norm(g, n->p, cnormDt, rtol, false)) {
      h->main.remove(n);
      temp.append(n);
    }
    n = next;
  }
  int iter = 1;
  while (temp.len > 0) {
    gather_list(temp);
    boris_list(temp);
    apply_its += temp.len;
    for (Node *n = temp.head; n;) {    /* stepNormTransfer(temp -> main), reverse */
      Node *next = n->next;
      if (step_norm(g, n->p, cnormDt, rtol, true)) {
        temp.remove(n);
        h->main.append(n);
      }
      n = next;
    }
    if (temp.len == 0) break;
    if (iter >= iter_max) break;
    iter += 1;
  }
  if (num_unconverged) *num_unconverged = temp.len;
  for (Node *n = temp.head; n;) {      /* :1678-1693: what is left goes back */
    Node *next = n->next;
    temp.remove(n);
    h->main.append(n);
    n = next;
  }
  for (Node *n = h->main.head; n; n = n->next)   /* setCurrentDensity: one deposit call per particle */
    if (orc_deposit_current(g, interp, 1, n->p.position, n->p.position_old, n->p.velocity, &n->p.weight, cnormDt, J))
      rc = -1;
  if (num_apply_its) *num_apply_its = apply_its;
  return rc;
}

/* back to SoA in id order (xbar, ubar) */
extern "C" void orc_aos_read(void *hv, double *x, double *v) {
  Aos *h = static_cast<Aos *>(hv);
  const long n = h->n;
  for (long p = 0; p < n; ++p) {
    const AosParticle &q = h->nodes[p]->p;
    for (int d = 0; d < h->D; ++d) x[d * n + p] = q.position[d];
    for (int c = 0; c < 3; ++c) v[c * n + p] = q.velocity[c];
  }
}
