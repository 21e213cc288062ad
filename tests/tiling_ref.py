"""numpy mirror of the tile plan of include/vgs.h (tile k -> rank k % world, slots, scatter/mirror),
used by the CPU tests to check the sharding logic without a GPU."""
import numpy as np


def plan(n, tile, world, symmetric):
    nt = -(-n // tile)
    n_tiles = nt * (nt + 1) // 2 if symmetric else nt * nt
    return nt, n_tiles, -(-n_tiles // world)


def tile_coords(k, nt, symmetric):
    if not symmetric:
        return k // nt, k % nt
    i, rem = 0, k
    while rem >= nt - i:
        rem -= nt - i
        i += 1
    return i, i + rem


def payload_of_rank(D, tile, rank, world, symmetric):
    n = D.shape[0]
    nt, n_tiles, slots = plan(n, tile, world, symmetric)
    out = np.zeros((slots, tile, tile), dtype=D.dtype)
    for q in range(slots):
        k = q * world + rank
        if k >= n_tiles:
            continue
        I, J = tile_coords(k, nt, symmetric)
        blk = D[I * tile:(I + 1) * tile, J * tile:(J + 1) * tile]
        out[q, :blk.shape[0], :blk.shape[1]] = blk
    return out


def assemble(gathered, n, tile, world, symmetric):
    nt, n_tiles, slots = plan(n, tile, world, symmetric)
    g = gathered.reshape(world, slots, tile, tile)
    D = np.zeros((n, n), dtype=gathered.dtype)
    for i in range(n):
        for j in range(n):
            I, J, ii, jj = i // tile, j // tile, i % tile, j % tile
            if symmetric:
                if I > J:
                    I, J, ii, jj = J, I, jj, ii
                k = I * nt - I * (I - 1) // 2 + (J - I)
            else:
                k = I * nt + J
            D[i, j] = g[k % world, k // world, ii, jj]
    return D
