"""Table reporters with the reference's on-disk formats (reporter.py:57-164, 564-632, 790-834), so that
``post_processing.py`` and any downstream tooling read the output of a GPU run unchanged:

  <out_path>/tab_iter.csv           "{iter:2d} {goal_close:.6f} {Fsm:.6f} {E_int:.4f} {J:.6f}"
  <out_path>/tab_iter_E.csv         "{iter:2d} {t_fs:.6f} {|E|:.6e}"            (not for plotting_flag tables_iter)
  <out_path>/<prop_id>/tables/tab_abs_{level}.csv / tab_real_{level}.csv   "{t_fs:.6f} {x:.6f} {value:.6e}"
  <out_path>/<prop_id>/tables/tab_tvals_{level}.csv    t <x> <x^2> <p> <p^2> E_n |norm| |overlp0| |overlpf| max|psi| max Re
  <out_path>/<prop_id>/tables/tab_tvals_fit.csv        t |E| freq_mult E_tot |overlp0_tot| |overlpf_tot| sigma_x,y,z

The HTML plot reporters (reporter.py:167-530, 635-788) are presentation only and are not reproduced;
``MultipleFitterReporter`` accepts their configuration and ignores it.
"""
from __future__ import annotations

import copy
import math
import os

from .config import ReportRootConfiguration

_OT = ReportRootConfiguration.ReportFitterConfiguration.OutputType


class PropagationReporter:
    def __init__(self, out_path: str, nlvls: int):
        self._out_path = out_path
        self._nlvls = nlvls

    def open(self):
        raise NotImplementedError()

    def close(self):
        raise NotImplementedError()

    def print_time_point_prop(self, l, psi, t, x, np, nt, moms, smoms, ener, norm, overlp0, overlpf, overlp_tot, ener_tot,
                              psi_max_abs, psi_max_real, E, freq_mult):
        raise NotImplementedError()


class TablePropagationReporter(PropagationReporter):
    def __init__(self, out_path, nlvls, conf):
        super().__init__(out_path, nlvls)
        self.conf = conf
        self.f_abs, self.f_real, self.f_prop, self.f_fit = [], [], [], None

    @staticmethod
    def name_template(name, level):
        return name.replace("{level}", str(level))

    def open(self):
        os.makedirs(self._out_path, exist_ok=True)
        mk = lambda tmpl, n: open(os.path.join(self._out_path, self.name_template(tmpl, n)), "w")
        self.f_abs = [mk(self.conf.tab_abs, n) for n in range(self._nlvls)]
        self.f_real = [mk(self.conf.tab_real, n) for n in range(self._nlvls)]
        self.f_prop = [mk(self.conf.tab_tvals, n) for n in range(self._nlvls)]
        self.f_fit = open(os.path.join(self._out_path, self.conf.tab_tvals_fit), "w")
        return self

    def close(self):
        for f in self.f_abs + self.f_real + self.f_prop + ([self.f_fit] if self.f_fit else []):
            f.close()
        self.f_abs, self.f_real, self.f_prop, self.f_fit = [], [], [], None

    def print_time_point_prop(self, l, psi, t, x, np, nt, moms, smoms, ener, norm, overlp0, overlpf, overlp_tot, ener_tot,
                              psi_max_abs, psi_max_real, E, freq_mult):
        if l % self.conf.mod_fileout != 0 or l < self.conf.lmin:
            return
        tf = t * 1e+15
        for n in range(self._nlvls):
            col = psi.f[n]
            self.f_abs[n].write("".join("{:.6f} {:.6f} {:.6e}\n".format(tf, x[i], abs(col[i])) for i in range(np)))
            self.f_real[n].write("".join("{:.6f} {:.6f} {:.6e}\n".format(tf, x[i], col[i].real) for i in range(np)))
            self.f_prop[n].write("{:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f}\n".format(
                tf, moms.x[n].real, moms.x2[n].real, moms.p[n].real, moms.p2[n].real, ener[n].real, abs(norm[n]),
                abs(overlp0[n]), abs(overlpf[n]), psi_max_abs[n], psi_max_real[n]))
            for f in (self.f_abs[n], self.f_real[n], self.f_prop[n]):
                f.flush()
        self.f_fit.write("{:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f}\n".format(
            tf, abs(E), freq_mult, ener_tot.real, abs(overlp_tot[0]), abs(overlp_tot[1]), smoms.x.real, smoms.y.real,
            smoms.z.real))
        self.f_fit.flush()


class MultiplePropagationReporter(PropagationReporter):
    def __init__(self, out_path, plotting_flag, nlvls, conf_rep_table, conf_rep_plot=None):
        super().__init__(out_path, nlvls)
        self.reps = []
        if conf_rep_table is not None and not conf_rep_table.is_empty() and plotting_flag in (_OT.ALL, _OT.TABLES):
            self.reps.append(TablePropagationReporter(os.path.join(out_path, "tables"), nlvls, conf_rep_table))

    @property
    def mod_fileout(self):          # lets the GPU path synchronise only on steps some member writes: gcd of the moduli
        g = 0
        for r in self.reps:
            g = math.gcd(g, int(r.conf.mod_fileout))
        return g if g > 0 else 1 << 30

    def open(self):
        for r in self.reps:
            r.open()
        return self

    def close(self):
        for r in self.reps:
            r.close()

    def print_time_point_prop(self, *a):
        for r in self.reps:
            r.print_time_point_prop(*a)


class FitterReporter:
    def open(self):
        raise NotImplementedError()

    def close(self):
        raise NotImplementedError()

    def print_iter_point_fitter(self, iter, goal_close, E_tlist, t_list, Fsm, E_int, J, nt):
        raise NotImplementedError()

    def create_propagation_reporter(self, prop_id, nlvls):
        raise NotImplementedError()


class TableFitterReporter(FitterReporter):
    def __init__(self, conf):
        self.conf = conf
        self.f_ifit = self.f_ifit_E = None

    def create_propagation_reporter(self, prop_id, nlvls):
        c = copy.deepcopy(self.conf)
        return TablePropagationReporter(os.path.join(c.out_path, prop_id), nlvls, c.propagation)

    def open(self):
        os.makedirs(self.conf.out_path, exist_ok=True)
        self.f_ifit = open(os.path.join(self.conf.out_path, self.conf.tab_iter), "w")
        if self.conf.plotting_flag != _OT.TABLES_ITER:
            self.f_ifit_E = open(os.path.join(self.conf.out_path, self.conf.tab_iter_E), "w")
        return self

    def close(self):
        for f in (self.f_ifit, self.f_ifit_E):
            if f:
                f.close()
        self.f_ifit = self.f_ifit_E = None

    def print_iter_point_fitter(self, iter, goal_close, E_tlist, t_list, Fsm, E_int, J, nt):
        if iter % self.conf.imod_fileout != 0 or iter < self.conf.imin:
            return
        self.f_ifit.write("{:2d} {:.6f} {:.6f} {:.4f} {:.6f}\n".format(int(iter), goal_close, Fsm.real, E_int, J))
        self.f_ifit.flush()
        if self.f_ifit_E is not None and E_tlist is not None:
            self.f_ifit_E.write("".join("{:2d} {:.6f} {:.6e}\n".format(iter, t_list[i] * 1e+15, abs(E_tlist[i]))
                                        for i in range(nt + 1)))
            self.f_ifit_E.flush()


class MultipleFitterReporter(FitterReporter):
    def __init__(self, conf_rep_table, conf_rep_plot=None):
        self.reps = []
        if not conf_rep_table.is_empty() and conf_rep_table.plotting_flag in (_OT.ALL, _OT.TABLES, _OT.TABLES_ITER):
            self.reps.append(TableFitterReporter(conf_rep_table))
        self.conf_rep_table, self.conf_rep_plot = conf_rep_table, conf_rep_plot

    def create_propagation_reporter(self, prop_id, nlvls):
        c = copy.deepcopy(self.conf_rep_table)
        prop_out = os.path.join(c.out_path, prop_id)
        if c.plotting_flag not in (_OT.TABLES_ITER, _OT.NONE):
            os.makedirs(prop_out, exist_ok=True)
        return MultiplePropagationReporter(prop_out, c.plotting_flag, nlvls, c.propagation, None)

    def open(self):
        for r in self.reps:
            r.open()
        return self

    def close(self):
        for r in self.reps:
            r.close()

    def print_iter_point_fitter(self, *a):
        for r in self.reps:
            r.print_iter_point_fitter(*a)
