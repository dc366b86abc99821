"""Back-of-the-envelope floors of the twelve conv_rows_tc launches of the bench step (R50 1024x320 bs32, thr 0.05), from
the constants measured on the B200: UTCHMMA cost per instruction, L2 -> SM feed rate, per-tile prologue /
epilogue clocks from scripts/tc_layer_trace.py, SM clock under load.  No GPU needed:  python scripts/conv_floor_model.py
Prints, per layer: tensor floor, feed floor, modelled time (floors + per-tile overheads) and the measured time."""
SMS, F_GHZ = 148, 1.5                      # SM clock under these launches: 1.35-1.55 GHz (CTA 0's clock span / event time)
FEED = 45.0                                # B/clk per SM when all SMs stream from L2 (tma_tile_rate: 690 clk per 32 KB)
GATHER4 = 33.0                             # B/clk per SM for gather4 loads (974 clk per 32 KB)

# (name, taps, c0, c1, cout, active rows, measured us in profiles/r02_bench_n1_final.json)
LAYERS = [("upconv(4,0)", 9, 2048, 0, 256, 10240, 405), ("upconv(4,1)", 9, 256, 1024, 256, 40960, 900),
          ("1x1 heads(4)", 1, 256, 0, 576, 40960, 134), ("taps(4)", 1, 576, 0, 63, 40960, 47),
          ("upconv(3,0)", 9, 256, 0, 128, 25658, 109), ("upconv(3,1)", 9, 128, 512, 128, 69892, 450),
          ("1x1 heads(3)", 1, 128, 0, 256, 69892, 80), ("taps(3)", 1, 256, 0, 54, 69892, 32),
          ("upconv(2,0)", 9, 128, 0, 64, 63307, 95), ("upconv(2,1)", 9, 64, 256, 64, 160314, 412),
          ("upconv(1,0)", 9, 64, 0, 32, 146838, 92), ("upconv(1,1)", 9, 32, 64, 32, 350128, 297)]


def tile_n(cout):
    return 128 if cout > 64 else (64 if cout > 32 else 32)


print("%-13s %5s %6s %7s | %8s %8s %8s | %8s %8s  %s" % ("layer", "N", "tiles", "chunks", "MMA us", "feed us", "ovh us", "model us", "meas us", "meas/MMA floor"))
tot = [0.0, 0.0, 0.0]
for name, taps, c0, c1, cout, rows, meas in LAYERS:
    bn = tile_n(cout)
    tiles = -(-rows // 256) * -(-cout // bn)
    chunks = taps * (-(-c0 // 32) + -(-c1 // 32))
    mma = 24 * (64 if bn == 128 else 52)                       # clk per chunk-tile, 3xTF32
    a_bytes = 256 * 128 * (1.25 / 3 if taps == 9 else 1.0)     # shared-tap stage: one fill per 3 chunks, ~25 % extras
    feed = a_bytes / (GATHER4 if taps == 9 else FEED) + bn * 256 / FEED
    # per-tile clocks outside the chunk loop (tables, first fill, last epoch + drain, epilogue), from the tile tracer
    ovh = {128: (7000 + 3000 + 1500 + 23000) if taps == 9 else (900 + 2600 + 2500 + 18000),
           64: (9000 + 2600 + 1000 + 7000) if taps == 9 else (1300 + 2600 + 1500 + 6000),
           32: (9000 + 2400 + 700 + 3500)}[bn]
    per_sm = tiles / SMS
    t_mma = per_sm * chunks * mma / F_GHZ / 1e3
    t_feed = per_sm * chunks * feed / F_GHZ / 1e3
    t_ovh = per_sm * ovh / F_GHZ / 1e3
    model = max(t_mma, t_feed) + t_ovh
    tot[0] += t_mma; tot[1] += model; tot[2] += meas
    print("%-13s %5d %6d %7d | %8.0f %8.0f %8.0f | %8.0f %8.0f  %.2f" % (name, bn, tiles, chunks, t_mma, t_feed, t_ovh, model, meas, meas / t_mma))
print("sum: MMA floor %.0f us, model %.0f us, measured %.0f us  (SM clock %.2f GHz; at the nominal 1.965 GHz the MMA floor would be %.0f us)" %
      (tot[0], tot[1], tot[2], F_GHZ, tot[0] * F_GHZ / 1.965))
