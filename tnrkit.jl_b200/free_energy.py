"""free_energy -- mirrors /root/reference/src/utility/free_energy.jl:14-21."""
import math


def free_energy(data, beta, scalefactor=2.0, initial_size=1.0):
    lnz = 0.0
    x = 1.0 - math.log(initial_size) / math.log(scalefactor)
    for i, z in enumerate(data, start=1):
        lnz += math.log(z) * scalefactor ** (x - i)
    return -lnz / beta
