// kessler.cpp -- oracle restatement of phys/module_mp_kessler.F:8-244 (called from
// phys/module_microphysics_driver.F:983-1003).  Test infrastructure only; see
// oracle.h header.
#include "oracle.h"
namespace orc {

void kessler(const Cx &x, A3 t, A3 qv, A3 qc, A3 qr, A3 rho, A3 pii, float dt_in, A3 z, A3 dz8w,
             A2 rainnc, A2 rainncv) {
  const int its = x.its, ite = x.ite, jts = x.jts, jte = x.jte, kts = x.kts, kte = x.kte;
  const float c1 = .001f, c2 = .001f, c3 = 2.2f, c4 = .875f;
  const float max_cr_sedimentation = 0.75f;
  const float dt = dt_in;
  const float f5 = svp2 * (svpt0 - svp3) * xlv / cp;
  const int nk = kte + 2;
  std::vector<float> vt(nk), prodk(nk), vtden(nk), rdzk(nk), rhok(nk), factor(nk), rdzw(nk);
  std::vector<float> prod_((size_t)(ite - its + 1) * nk);
  auto prod = [&](int i, int k) -> float & { return prod_[(size_t)(i - its) * nk + k]; };

  for (int j = jts; j <= jte; j++) {
    for (int i = its; i <= ite; i++) {
      float crmax = 0.f;
      for (int k = 1; k <= kte; k++) {
        prodk[k] = qr(i, k, j);
        rhok[k] = rho(i, k, j);
        float qrr = std::max(0.f, qr(i, k, j) * 0.001f * rhok[k]);
        vtden[k] = std::sqrt(rhok[1] / rhok[k]);
        vt[k] = 36.34f * std::pow(qrr, 0.1364f) * vtden[k];
        rdzw[k] = 1.f / dz8w(i, k, j);
        crmax = std::max(vt[k] * dt * rdzw[k], crmax);
      }
      for (int k = 1; k <= kte - 1; k++) rdzk[k] = 1.f / (z(i, k + 1, j) - z(i, k, j));
      rdzk[kte] = 1.f / (z(i, kte, j) - z(i, kte - 1, j));
      int nfall = std::max(1, (int)std::lround(0.5f + crmax / max_cr_sedimentation));
      float dtfall = dt / (float)nfall;
      float time_sediment = dt;
      while (nfall > 0) {
        time_sediment = time_sediment - dtfall;
        for (int k = 1; k <= kte - 1; k++) factor[k] = dtfall * rdzk[k] / rhok[k];
        factor[kte] = dtfall * rdzk[kte];
        float ppt = 0.f;
        int k = 1;
        ppt = rhok[k] * prodk[k] * vt[k] * dtfall / rhowater;
        rainncv(i, j) = ppt * 1000.f;
        rainnc(i, j) = rainnc(i, j) + ppt * 1000.f;
        for (k = kts; k <= kte - 1; k++)
          prodk[k] = prodk[k] - factor[k] * (rhok[k] * prodk[k] * vt[k] - rhok[k + 1] * prodk[k + 1] * vt[k + 1]);
        k = kte;
        prodk[k] = prodk[k] - factor[k] * prodk[k] * vt[k];
        if (nfall > 1) {
          nfall = nfall - 1;
          crmax = 0.f;
          for (k = kts; k <= kte; k++) {
            float qrr = std::max(0.f, prodk[k] * 0.001f * rhok[k]);
            vt[k] = 36.34f * std::pow(qrr, 0.1364f) * vtden[k];
            crmax = std::max(vt[k] * time_sediment * rdzw[k], crmax);
          }
          int nfall_new = std::max(1, (int)std::lround(0.5f + crmax / max_cr_sedimentation));
          if (nfall_new != nfall) {
            nfall = nfall_new;
            dtfall = time_sediment / (float)nfall;
          }
        } else {
          for (k = kts; k <= kte; k++) prod(i, k) = prodk[k];
          nfall = 0;
        }
      }
    }
    for (int k = kts; k <= kte; k++)
      for (int i = its; i <= ite; i++) {
        float factorn = 1.0f / (1.f + c3 * dt * std::pow(std::max(0.f, qr(i, k, j)), c4));
        float qrprod = qc(i, k, j) * (1.0f - factorn) + factorn * c1 * dt * std::max(qc(i, k, j) - c2, 0.f);
        float rcgs = 0.001f * rho(i, k, j);
        qc(i, k, j) = std::max(qc(i, k, j) - qrprod, 0.f);
        qr(i, k, j) = (qr(i, k, j) + prod(i, k) - qr(i, k, j));
        qr(i, k, j) = std::max(qr(i, k, j) + qrprod, 0.f);
        float temp = pii(i, k, j) * t(i, k, j);
        float pressure = 1.000e+05f * std::pow(pii(i, k, j), 1004.f / 287.f);
        float gam = 2.5e+06f / (1004.f * pii(i, k, j));
        float es = 1000.f * svp1 * std::exp(svp2 * (temp - svpt0) / (temp - svp3));
        float qvs = ep_2 * es / (pressure - es);
        float pr = (qv(i, k, j) - qvs) / (1.f + pressure / (pressure - es) * qvs * f5 / ((temp - svp3) * (temp - svp3)));
        prod(i, k) = pr;
        float dimv = std::max(qvs - qv(i, k, j), 0.f);  // DIM(qvs,qv)
        float ern = std::min(std::min(dt * (((1.6f + 124.9f * std::pow(rcgs * qr(i, k, j), .2046f)) * std::pow(rcgs * qr(i, k, j), .525f)) /
                                            (2.55e8f / (pressure * qvs) + 5.4e5f)) *
                                          (dimv / (rcgs * qvs)),
                                      std::max(-pr - qc(i, k, j), 0.f)),
                             qr(i, k, j));
        float product = std::max(pr, -qc(i, k, j));
        t(i, k, j) = t(i, k, j) + gam * (product - ern);
        qv(i, k, j) = std::max(qv(i, k, j) - product + ern, 0.f);
        qc(i, k, j) = qc(i, k, j) + product;
        qr(i, k, j) = qr(i, k, j) - ern;
      }
  }
}

}  // namespace orc
