"""Unit constants, numerically identical to the reference's ``ndsimulator/constant.py:2-15``."""
#: Boltzmann constant in eV/K
kB = 8.6173303e-5
#: electron volt in J
eV = 1.60217733e-19
#: angstrom in m
angstrom = 1.0e-10
#: femtosecond in s
fs = 1e-15
#: kcal/mol in eV
kcal = 0.04337209302
#: atomic mass unit in eV, angstrom, fs units
au_mass = 1.67e-27 / (eV / angstrom * fs / angstrom * fs)
#: scaling of the Mueller-Brown potential
escale = 0.2
lscale = 16.0
