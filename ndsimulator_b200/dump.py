"""Reference-format text dumps (``pos.dat force.dat velocity.dat colvar.dat thermo.dat fix{i}.dat``).

Replaces ``ndsimulator/io/dump.py:52-102`` for selected walkers: the rows come from frames recorded on
the device by the fused kernel (``nds_drain_frames``), formatted like the reference's
``" ".join(f"{x}" ...)``.  Quirks kept:

* the thermo row printed beside a dump row is the stat sample taken BEFORE it (``Dump.write`` runs before
  ``Stat.append`` in the loop, ``ndrun.py:247-251``): the kernel records that sample as its own frame;
* 1-output colvars of ``two.py`` / ``five.py`` return shape (1, 1), so ``colvar.dat`` shows ``[value]``;
* ``fix{i}.dat`` of a metadynamics bias lists the hills deposited since the previous dump as
  ``idx c... sigma... w`` followed by a blank line (``gaus.py:440-461``); an umbrella bias prints
  ``None`` (``umb.py:75-76`` returns nothing).

Walker 0 writes into ``<root>/<run_name>/`` exactly like the reference; walker k > 0 into
``<root>/<run_name>/walker_<k>/``.
"""
import os

import numpy as np

_INT_JACOBIAN_CVS = {"ExtractX", "ExtractY", "XmY", "XpY", "Proj5dto1dx0"}  # integer jacob arrays (two.py, five.py:7)
_BRACKET_CVS = {"ExtractX", "ExtractY", "XmY", "XpY", "Proj5dto1dx0", "Proj5dto1dx2", "Proj5dto1d"}


def _fmt(v):
    return repr(float(v))


class DumpWriter:
    def __init__(self, spec, n_walkers):
        self.spec = spec
        self.n = n_walkers
        self.dirs = []
        base = os.path.join(spec.root, spec.run_name)
        for k in range(n_walkers):
            d = base if k == 0 else os.path.join(base, f"walker_{k}")
            os.makedirs(d, exist_ok=True)
            self.dirs.append(d)
        self.files = []
        self.fix_names = []
        for b in spec.biases:
            if b == "umb":
                self.fix_names += ["umb"] * max(len(spec.umb_k), 1)
            else:
                self.fix_names.append(b)
        self.lag = None          # last stat sample seen (per walker rows)
        self.hills_dumped = 0
        self.umb_vr = None

    def open_files(self, append=False):
        """append: continue the files of an interrupted run (NDRun.load_checkpoint); no second header."""
        names = ["thermo", "pos", "force", "velocity", "colvar"] + [f"fix{i}" for i in range(len(self.fix_names))]
        for d in self.dirs:
            self.files.append({n: open(os.path.join(d, f"{n}.dat"), "a" if append else "w") for n in names})
        header = "#time dt T pe ke biase totale" + "".join(f" fix_VR_{i}" for i in range(len(self.fix_names)))
        for f in self.files:
            if not (append and f["thermo"].tell() > 0):
                f["thermo"].write(header + "\n")  # dump.py:71-74

    def close_files(self):
        for f in self.files:
            for h in f.values():
                h.close()
        self.files = []

    def write_frames(self, frames, names, hills=None, hills_at=None):
        """frames: [n][width][walkers] as returned by Engine.drain_frames(); hills = (centers, heights
        [, per-hill variances of an adaptive table]) of the current table (walker 0 / shared) for fix{i}.dat; hills_at(step) = number of hills that
        existed when the dump row of that step was written."""
        spec = self.spec
        nd, cd = spec.ndim, spec.colvardim
        idx = {n: i for i, n in enumerate(names)}
        for fr in frames:
            step = int(fr[idx["step"], 0])
            is_dump = step % spec.dump_freq == 0
            if is_dump:
                lag = fr if (self.lag is None or step == 0) else self.lag
                for w, f in enumerate(self.files):
                    if np.isnan(fr[idx["x0"], w]):
                        continue  # walker stopped (committor) before this frame
                    f["pos"].write(" ".join(_fmt(fr[idx[f"x{d}"], w]) for d in range(nd)) + "\n")
                    f["force"].write(" ".join(_fmt(fr[idx[f"f{d}"], w]) for d in range(nd)) + "\n")
                    f["velocity"].write(" ".join(_fmt(fr[idx[f"v{d}"], w]) for d in range(nd)) + "\n")
                    if spec.colvar in _BRACKET_CVS:
                        # f"{c}" of a (1,)-shaped ndarray: numpy's 8-digit array formatting (A28)
                        f["colvar"].write(" ".join(str(np.array([fr[idx[f"c{k}"], w]])) for k in range(cd)) + "\n")
                    else:
                        f["colvar"].write(" ".join(_fmt(fr[idx[f"c{k}"], w]) for k in range(cd)) + "\n")
                    # NDRun.time starts as the int 0 (ndrun.py:200); atoms.biase stays the int 0 without
                    # biases (md.py:213): both print as "0" in the reference's files
                    t_txt = "0" if step == 0 else _fmt(lag[idx["time"], w])
                    b_txt = "0" if not spec.biases else _fmt(lag[idx["biase"], w])
                    line = " ".join([t_txt, _fmt(spec.dt), _fmt(lag[idx["T"], w]), _fmt(lag[idx["pe"], w]),
                                     _fmt(lag[idx["ke"], w]), b_txt, _fmt(lag[idx["totale"], w])])
                    n_umb_seen = 0
                    for name in self.fix_names:  # dump.py:98: current fix.VR of every bias
                        if name == "mtd":
                            line += f" {_fmt(fr[idx['VR'], w])}"
                        elif name == "umb":
                            if f"VRu{n_umb_seen}" in idx:  # refreshed by every deposition (gaus.py:138-146)
                                line += f" {_fmt(fr[idx[f'VRu{n_umb_seen}'], w])}"
                            else:                          # umbrella only: Harmonic.VR stays at its begin() value
                                line += f" {_fmt(self.umb_vr[w]) if self.umb_vr is not None else 0.0}"
                            n_umb_seen += 1
                        else:
                            line += " 0"
                    f["thermo"].write(line + "\n")
                    for i, name in enumerate(self.fix_names):
                        if name == "mtd" and hills is not None and w == 0:
                            c, h = hills[0], hills[1]
                            s2 = hills[2] if len(hills) > 2 else None  # adaptive: _sigma_history[idx].flat
                            out = ""
                            upto = len(h) if hills_at is None else min(len(h), hills_at(step))
                            for j in range(self.hills_dumped, upto):
                                if s2 is None:
                                    widths = " ".join(_fmt(v) for v in spec.mtd_sigma)
                                elif j == 0 and spec.colvar in _INT_JACOBIAN_CVS:
                                    # first hill: J.J^T of an INTEGER Jacobian array prints as an int (two.py:8,19,30,41)
                                    widths = str(int(round(s2[j])))
                                else:
                                    widths = _fmt(s2[j])
                                out += f"{j} " + " ".join(_fmt(v) for v in c[j]) + " " + widths + " " + _fmt(h[j]) + "\n"
                            f[f"fix{i}"].write(out + "\n")
                        elif name == "umb":
                            f[f"fix{i}"].write("None\n")
                        else:
                            f[f"fix{i}"].write("\n")
                if hills is not None:
                    self.hills_dumped = len(hills[1]) if hills_at is None else min(len(hills[1]), hills_at(step))
            if step % spec.stat_freq == 0:
                self.lag = fr
        for f in self.files:
            for h in f.values():
                h.flush()
