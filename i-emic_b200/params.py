"""The 30 continuation parameters of THCM: XML names <-> Fortran index.

Mirrors THCM::par2int / int2par (src/ocean/THCM.C:1841-1942) and par.F90:38-67.
"""
PAR_NAMES = {
    0: "Time", 1: "AL_T", 2: "Rayleigh-Number", 3: "Vertical Ekman-Number", 4: "Horizontal Ekman-Number",
    5: "Rossby-Number", 6: "MIXP", 7: "RESC", 8: "SPL1", 9: "Salinity Homotopy", 10: "Solar Forcing",
    11: "Horizontal Peclet-Number", 12: "Vertical Peclet-Number", 13: "P_VC", 14: "LAMB",
    15: "Salinity Forcing", 16: "Wind Forcing", 17: "Temperature Forcing", 18: "Nonlinear Factor",
    19: "Combined Forcing", 20: "ARCL", 21: "NLES", 22: "IFRICB", 23: "CONT", 24: "Energy",
    25: "ALPC", 26: "CMPR", 27: "Flux Perturbation", 28: "Salinity Perturbation", 29: "MKAP", 30: "SPL2",
}
# Fortran enumeration names (par.F90:38-67)
PAR_INDEX = dict(AL_T=1, RAYL=2, EK_V=3, EK_H=4, ROSB=5, MIXP=6, RESC=7, SPL1=8, HMTP=9, SUNP=10, PE_H=11,
                 PE_V=12, P_VC=13, LAMB=14, SALT=15, WIND=16, TEMP=17, BIOT=18, COMB=19, ARCL=20, NLES=21,
                 IFRICB=22, CONT=23, ENER=24, ALPC=25, CMPR=26, FPER=27, SPER=28, MKAP=29, SPL2=30)


def par_index(name):
    """Accepts a Fortran enumeration name ('COMB'), an XML name ('Combined Forcing') or an int."""
    if isinstance(name, int):
        return name
    if name in PAR_INDEX:
        return PAR_INDEX[name]
    for k, v in PAR_NAMES.items():
        if v == name:
            return k
    raise KeyError(f"unknown THCM parameter {name!r}")
