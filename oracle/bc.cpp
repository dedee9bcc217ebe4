// bc.cpp -- oracle restatement of share/module_bc.F set_physical_bc2d (:202-648)
// and set_physical_bc3d (:651-1113): periodic / symmetric / open ghost-zone
// fills (specified, nested, polar are out of scope).  bdyzone = 4 (:33).
// Test infrastructure only; see oracle.h header.
#include "oracle.h"
namespace orc {

static const int bdyzone = 4;

void set_physical_bc3d(const Cx &x, A3 dat, char variable_in) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde, kde = x.kde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  char variable = variable_in;
  if (variable >= 'A' && variable <= 'Z') variable = (char)(variable - 'A' + 'a');
  bool open_bc_copy = false;
  if (variable == 'u' || variable == 'v' || variable == 'w' || variable == 't' || variable == 'd' || variable == 'e' ||
      variable == 'x' || variable == 'y' || variable == 'f' || variable == 'r' || variable == 'p')
    open_bc_copy = true;
  int istag = -1, jstag = -1;
  int k_end = std::max(1, std::min(kde - 1, kte));
  if (variable == 'u' || variable == 'x') istag = 0;
  if (variable == 'v' || variable == 'y') jstag = 0;
  if (variable == 'd') { istag = 0; jstag = 0; }
  if (variable == 'e') { istag = 0; k_end = std::min(kde, kte); }
  if (variable == 'f') { jstag = 0; k_end = std::min(kde, kte); }
  if (variable == 'w') k_end = std::min(kde, kte);

  if (x.c.periodic_x) {
    if (ids == x.ips && ide == x.ipe) {
      if (its == ids)
        for (int j = std::max(jds, jts - 1); j <= std::min(jte + 1, jde + jstag); j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = 0; i >= -(bdyzone - 1); i--) dat(ids + i - 1, k, j) = dat(ide + i - 1, k, j);
      if (ite == ide)
        for (int j = std::max(jds, jts - 1); j <= std::min(jte + 1, jde + jstag); j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = -istag; i <= bdyzone; i++) dat(ide + i + istag, k, j) = dat(ids + i + istag, k, j);
    }
  } else {
    if (x.c.symmetric_xs && its == ids) {
      const int j0 = std::max(jds, jts - 1), j1 = std::min(jte + 1, jde + jstag);
      if (istag == -1) {
        for (int j = j0; j <= j1; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = 1; i <= bdyzone; i++) dat(ids - i, k, j) = dat(ids + i - 1, k, j);
      } else if (variable == 'u') {
        for (int j = j0; j <= j1; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = 1; i <= bdyzone; i++) dat(ids - i, k, j) = -dat(ids + i, k, j);
      } else {
        for (int j = j0; j <= j1; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = 1; i <= bdyzone; i++) dat(ids - i, k, j) = dat(ids + i, k, j);
      }
    }
    if (x.c.symmetric_xe && ite == ide) {
      const int j0 = std::max(jds, jts - 1), j1 = std::min(jte + 1, jde + jstag);
      if (istag == -1) {
        for (int j = j0; j <= j1; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = 1; i <= bdyzone; i++) dat(ide + i - 1, k, j) = dat(ide - i, k, j);
      } else if (variable == 'u') {
        for (int j = j0; j <= j1; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = 1; i <= bdyzone; i++) dat(ide + i, k, j) = -dat(ide - i, k, j);
      } else {
        for (int j = j0; j <= j1; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = 1; i <= bdyzone; i++) dat(ide + i, k, j) = dat(ide - i, k, j);
      }
    }
    if (x.oxs() && its == ids && open_bc_copy) {
      for (int j = jts - bdyzone; j <= std::min(jte, jde + jstag) + bdyzone; j++)
        for (int k = kts; k <= k_end; k++) {
          dat(ids - 1, k, j) = dat(ids, k, j);
          dat(ids - 2, k, j) = dat(ids, k, j);
          dat(ids - 3, k, j) = dat(ids, k, j);
        }
    }
    if (x.oxe() && ite == ide && open_bc_copy) {
      if (variable != 'u' && variable != 'x') {
        for (int j = jts - bdyzone; j <= std::min(jte, jde + jstag) + bdyzone; j++)
          for (int k = kts; k <= k_end; k++) {
            dat(ide, k, j) = dat(ide - 1, k, j);
            dat(ide + 1, k, j) = dat(ide - 1, k, j);
            dat(ide + 2, k, j) = dat(ide - 1, k, j);
          }
      } else {
        for (int j = std::max(jds, jts - 1) - bdyzone; j <= std::min(jte + 1, jde + jstag) + bdyzone; j++)
          for (int k = kts; k <= k_end; k++) {
            dat(ide + 1, k, j) = dat(ide, k, j);
            dat(ide + 2, k, j) = dat(ide, k, j);
            dat(ide + 3, k, j) = dat(ide, k, j);
          }
      }
    }
  }

  int i_start = std::max(ids, its - 1), i_end = std::min(ite + 1, ide + istag);
  if (its == ids) i_start = x.ims;
  if (ite == ide) i_end = x.ime;

  if (x.c.periodic_y) {
    if (jds == x.jps && jde == x.jpe) {
      if (jts == jds)
        for (int j = 0; j >= -(bdyzone - 1); j--)
          for (int k = kts; k <= k_end; k++)
            for (int i = i_start; i <= i_end; i++) dat(i, k, jds + j - 1) = dat(i, k, jde + j - 1);
      if (jte == jde)
        for (int j = -jstag; j <= bdyzone; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = i_start; i <= i_end; i++) dat(i, k, jde + j + jstag) = dat(i, k, jds + j + jstag);
    }
  } else {
    if (x.c.symmetric_ys && jts == jds) {
      if (jstag == -1) {
        for (int j = 1; j <= bdyzone; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = i_start; i <= i_end; i++) dat(i, k, jds - j) = dat(i, k, jds + j - 1);
      } else if (variable == 'v') {
        for (int j = 1; j <= bdyzone; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = i_start; i <= i_end; i++) dat(i, k, jds - j) = -dat(i, k, jds + j);
      } else {
        for (int j = 1; j <= bdyzone; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = i_start; i <= i_end; i++) dat(i, k, jds - j) = dat(i, k, jds + j);
      }
    }
    if (x.c.symmetric_ye && jte == jde) {
      if (jstag == -1) {
        for (int j = 1; j <= bdyzone; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = i_start; i <= i_end; i++) dat(i, k, jde + j - 1) = dat(i, k, jde - j);
      } else if (variable == 'v') {
        for (int j = 1; j <= bdyzone; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = i_start; i <= i_end; i++) dat(i, k, jde + j) = -dat(i, k, jde - j);
      } else {
        for (int j = 1; j <= bdyzone; j++)
          for (int k = kts; k <= k_end; k++)
            for (int i = i_start; i <= i_end; i++) dat(i, k, jde + j) = dat(i, k, jde - j);
      }
    }
    if (x.oys() && jts == jds && open_bc_copy) {
      for (int k = kts; k <= k_end; k++)
        for (int i = i_start; i <= i_end; i++) {
          dat(i, k, jds - 1) = dat(i, k, jds);
          dat(i, k, jds - 2) = dat(i, k, jds);
          dat(i, k, jds - 3) = dat(i, k, jds);
        }
    }
    if (x.oye() && jte == jde && open_bc_copy) {
      if (variable != 'v' && variable != 'y') {
        for (int k = kts; k <= k_end; k++)
          for (int i = i_start; i <= i_end; i++) {
            dat(i, k, jde) = dat(i, k, jde - 1);
            dat(i, k, jde + 1) = dat(i, k, jde - 1);
            dat(i, k, jde + 2) = dat(i, k, jde - 1);
          }
      } else {
        for (int k = kts; k <= k_end; k++)
          for (int i = i_start; i <= i_end; i++) {
            dat(i, k, jde + 1) = dat(i, k, jde);
            dat(i, k, jde + 2) = dat(i, k, jde);
            dat(i, k, jde + 3) = dat(i, k, jde);
          }
      }
    }
  }
}

void set_physical_bc2d(const Cx &x, A2 dat, char variable_in) {
  const int ids = x.ids, ide = x.ide, jds = x.jds, jde = x.jde;
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte;
  char variable = variable_in;
  if (variable >= 'A' && variable <= 'Z') variable = (char)(variable - 'A' + 'a');
  bool open_bc_copy = false;
  if (variable == 'u' || variable == 'v' || variable == 'w' || variable == 't' || variable == 'x' || variable == 'y' ||
      variable == 'r' || variable == 'p')
    open_bc_copy = true;
  int istag = -1, jstag = -1;
  if (variable == 'u' || variable == 'x') istag = 0;
  if (variable == 'v' || variable == 'y') jstag = 0;
  if (variable == 'd') { istag = 0; jstag = 0; }
  if (variable == 'e') istag = 0;
  if (variable == 'f') jstag = 0;

  if (x.c.periodic_x) {
    if (ids == x.ips && ide == x.ipe) {
      if (its == ids)
        for (int j = std::max(jds, jts - 1); j <= std::min(jte + 1, jde + jstag); j++)
          for (int i = 0; i >= -(bdyzone - 1); i--) dat(ids + i - 1, j) = dat(ide + i - 1, j);
      if (ite == ide)
        for (int j = std::max(jds, jts - 1); j <= std::min(jte + 1, jde + jstag); j++)
          for (int i = -istag; i <= bdyzone; i++) dat(ide + i + istag, j) = dat(ids + i + istag, j);
    }
  } else {
    const int j0 = std::max(jds, jts - 1), j1 = std::min(jte + 1, jde + jstag);
    if (x.c.symmetric_xs && its == ids) {
      if (variable != 'u' && variable != 'x') {
        for (int j = j0; j <= j1; j++)
          for (int i = 1; i <= bdyzone; i++) dat(ids - i, j) = dat(ids + i - 1, j);
      } else if (variable == 'u') {
        for (int j = j0; j <= j1; j++)
          for (int i = 0; i <= bdyzone - 1; i++) dat(ids - i, j) = -dat(ids + i, j);
      } else {
        for (int j = j0; j <= j1; j++)
          for (int i = 0; i <= bdyzone - 1; i++) dat(ids - i, j) = dat(ids + i, j);
      }
    }
    if (x.c.symmetric_xe && ite == ide) {
      if (variable != 'u' && variable != 'x') {
        for (int j = j0; j <= j1; j++)
          for (int i = 1; i <= bdyzone; i++) dat(ide + i - 1, j) = dat(ide - i, j);
      } else if (variable == 'u') {
        for (int j = j0; j <= j1; j++)
          for (int i = 0; i <= bdyzone - 1; i++) dat(ide + i, j) = -dat(ide - i, j);
      } else {
        for (int j = j0; j <= j1; j++)
          for (int i = 0; i <= bdyzone - 1; i++) dat(ide + i, j) = dat(ide - i, j);
      }
    }
    if (x.oxs() && its == ids && open_bc_copy) {
      for (int j = j0; j <= j1; j++) {
        dat(ids - 1, j) = dat(ids, j);
        dat(ids - 2, j) = dat(ids, j);
        dat(ids - 3, j) = dat(ids, j);
      }
    }
    if (x.oxe() && ite == ide && open_bc_copy) {
      if (variable != 'u' && variable != 'x') {
        for (int j = j0; j <= j1; j++) {
          dat(ide, j) = dat(ide - 1, j);
          dat(ide + 1, j) = dat(ide - 1, j);
          dat(ide + 2, j) = dat(ide - 1, j);
        }
      } else {
        for (int j = j0; j <= j1; j++) {
          dat(ide + 1, j) = dat(ide, j);
          dat(ide + 2, j) = dat(ide, j);
          dat(ide + 3, j) = dat(ide, j);
        }
      }
    }
  }

  int istart = std::max(ids, its - 1), iend = std::min(ite + 1, ide + istag);
  if (its == ids) istart = x.ims;
  if (ite == ide) iend = x.ime;

  if (x.c.periodic_y) {
    if (jds == x.jps && jde == x.jpe) {
      if (jts == jds)
        for (int j = 0; j >= -(bdyzone - 1); j--)
          for (int i = istart; i <= iend; i++) dat(i, jds + j - 1) = dat(i, jde + j - 1);
      if (jte == jde)
        for (int j = -jstag; j <= bdyzone; j++)
          for (int i = istart; i <= iend; i++) dat(i, jde + j + jstag) = dat(i, jds + j + jstag);
    }
  } else {
    if (x.c.symmetric_ys && jts == jds) {
      if (variable != 'v' && variable != 'y') {
        for (int j = 1; j <= bdyzone; j++)
          for (int i = istart; i <= iend; i++) dat(i, jds - j) = dat(i, jds + j - 1);
      } else if (variable == 'v') {
        for (int j = 1; j <= bdyzone; j++)
          for (int i = istart; i <= iend; i++) dat(i, jds - j) = -dat(i, jds + j);
      } else {
        for (int j = 1; j <= bdyzone; j++)
          for (int i = istart; i <= iend; i++) dat(i, jds - j) = dat(i, jds + j);
      }
    }
    if (x.c.symmetric_ye && jte == jde) {
      if (variable != 'v' && variable != 'y') {
        for (int j = 1; j <= bdyzone; j++)
          for (int i = istart; i <= iend; i++) dat(i, jde + j - 1) = dat(i, jde - j);
      } else if (variable == 'v') {
        for (int j = 1; j <= bdyzone; j++)
          for (int i = istart; i <= iend; i++) dat(i, jde + j) = -dat(i, jde - j);
      } else {
        for (int j = 1; j <= bdyzone; j++)
          for (int i = istart; i <= iend; i++) dat(i, jde + j) = dat(i, jde - j);
      }
    }
    if (x.oys() && jts == jds && open_bc_copy) {
      for (int i = istart; i <= iend; i++) {
        dat(i, jds - 1) = dat(i, jds);
        dat(i, jds - 2) = dat(i, jds);
        dat(i, jds - 3) = dat(i, jds);
      }
    }
    if (x.oye() && jte == jde && open_bc_copy) {
      if (variable != 'v' && variable != 'y') {
        for (int i = istart; i <= iend; i++) {
          dat(i, jde) = dat(i, jde - 1);
          dat(i, jde + 1) = dat(i, jde - 1);
          dat(i, jde + 2) = dat(i, jde - 1);
        }
      } else {
        for (int i = istart; i <= iend; i++) {
          dat(i, jde + 1) = dat(i, jde);
          dat(i, jde + 2) = dat(i, jde);
          dat(i, jde + 3) = dat(i, jde);
        }
      }
    }
  }
}

}  // namespace orc
