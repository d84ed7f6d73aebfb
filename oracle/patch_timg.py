#!/usr/bin/env python
"""Build-time patch of a COPY of forward/sv_curv_col_el.c (TEST INFRASTRUCTURE; the reference tree is never modified and no
reference source enters the repository: the output goes to oracle/_ref/build/, which is git-ignored).

sv_curv_col_el_rhs_timg_z2 reads `veczt[n_free-(n-n_free)]` with a negative index at the free-surface row whenever the zeta
operator is {-1,0,1,2,3} (forward/sv_curv_col_el.c:162, 243, 290; SURVEY.md section 8c hazard 1). The three occurrences are replaced by

  zero   : (idx >= 0 ? veczt[idx] : 0.0f)                         -- the deterministic oracle of the parity default (CGFD_TIMG_ZERO)
  mirror : (idx >= 0 ? veczt[idx] : flux at grid row k + fdz_indx[n] - 2 (n - n_free))
           -- the in-source commented alternative (forward/sv_curv_col_el.c:154-159) with the same 2 Tsrc - image form as line 162;
              identical to line 162 wherever 162 is defined (CGFD_TIMG_MIRROR)

usage: patch_timg.py <zero|mirror> <in.c> <out.c>
"""
import sys

NEEDLE = "veczt[n_free-(n-n_free)]"
# stress triplets of the three momentum components, in the order the blocks appear in the function (hVx, hVy, hVz)
TRIPLETS = (("Txx", "Txy", "Txz"), ("Txy", "Tyy", "Tyz"), ("Txz", "Tyz", "Tzz"))

HEADER = """/* PATCHED COPY (oracle/patch_timg.py %s) of forward/sv_curv_col_el.c -- test infrastructure, not product */
#define CGFD_VECZT_ZERO(idx) ((idx) >= 0 ? veczt[(idx)] : 0.0f)
#define CGFD_VECZT_MIRROR(idx, T1, T2, T3) ((idx) >= 0 ? veczt[(idx)] : \\
   jac3d[iptr + (fdz_indx[n] - 2*(n-n_free)) * siz_slice] * ( \\
     zt_x[iptr + (fdz_indx[n] - 2*(n-n_free)) * siz_slice] * T1[iptr + (fdz_indx[n] - 2*(n-n_free)) * siz_slice] + \\
     zt_y[iptr + (fdz_indx[n] - 2*(n-n_free)) * siz_slice] * T2[iptr + (fdz_indx[n] - 2*(n-n_free)) * siz_slice] + \\
     zt_z[iptr + (fdz_indx[n] - 2*(n-n_free)) * siz_slice] * T3[iptr + (fdz_indx[n] - 2*(n-n_free)) * siz_slice] ))
"""


def main():
    mode, src, dst = sys.argv[1:4]
    text = open(src).read()
    parts = text.split(NEEDLE)
    if len(parts) != 4:
        sys.exit("patch_timg: expected 3 occurrences of %s, found %d" % (NEEDLE, len(parts) - 1))
    out = [parts[0]]
    for n in range(3):
        if mode == "zero":
            out.append("CGFD_VECZT_ZERO(n_free-(n-n_free))")
        elif mode == "mirror":
            out.append("CGFD_VECZT_MIRROR(n_free-(n-n_free), %s, %s, %s)" % TRIPLETS[n])
        else:
            sys.exit("patch_timg: mode must be zero or mirror")
        out.append(parts[n + 1])
    open(dst, "w").write(HEADER % mode + "".join(out))
    print("patch_timg: %s -> %s (%s, 3 occurrences)" % (src, dst, mode))


if __name__ == "__main__":
    main()
