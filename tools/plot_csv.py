"""Regenerates the reference's figures (plots/*_gemv_flops, *_gemv_error_u-1,1,
*_dot_flops, *_dot_error_median_u-1,1 -- the reference ships the pictures for
A100/V100 but neither data nor script) from the CSV the drivers print, as plain
SVG (no matplotlib in this image).

    python tools/plot_csv.py gemv_flops  gemv.csv        plots/b200_gemv_flops.svg
    python tools/plot_csv.py gemv_error  gemv_error.csv  plots/b200_gemv_error_u-1,1.svg
    python tools/plot_csv.py dot_flops   dot.csv         plots/b200_dot_flops.svg
    python tools/plot_csv.py dot_error   dot_error.csv   plots/b200_dot_error_median_u-1,1.svg
    python tools/plot_csv.py trsv_time   trsv.csv        plots/b200_trsv_time.svg

GFLOP/s: GEMV 2 n^2 / t, DOT 2 n / t (t = the driver's minimum of ten, in ms).
"""
import math
import sys

COLORS = ["#a2142f", "#77ac30", "#edb120", "#0072bd", "#4dbeee", "#d95319", "#7e2f8e", "#17becf",
          "#e377c2", "#7f7f7f"]


def read_csv(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("---"):
                break
            rows.append(line.split(";"))
    header = rows[0]
    data = [[float(v) for v in r] for r in rows[1:] if len(r) == len(header)]
    return header, data


def nice_ticks(lo, hi, count=6):
    span = hi - lo
    step = 10 ** math.floor(math.log10(span / count))
    for mult in (1, 2, 5, 10):
        if span / (step * mult) <= count:
            step *= mult
            break
    first = math.ceil(lo / step) * step
    ticks = []
    t = first
    while t <= hi + 1e-9 * span:
        ticks.append(t)
        t += step
    return ticks


def svg_plot(series, xlabel, ylabel, title, out, logy=False):
    W, H, L, R, T, B = 900, 560, 90, 250, 40, 60
    xs = [x for _, pts in series for x, _ in pts]
    ys = [y for _, pts in series for _, y in pts if (y > 0 or not logy)]
    x0, x1 = min(xs), max(xs)
    if logy:
        y0, y1 = math.floor(math.log10(min(ys))), math.ceil(math.log10(max(ys)))
    else:
        y0, y1 = 0.0, max(ys) * 1.05

    def px(x):
        return L + (x - x0) / (x1 - x0) * (W - L - R)

    def py(y):
        v = math.log10(y) if logy else y
        return H - B - (v - y0) / (y1 - y0) * (H - T - B)

    o = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="sans-serif" '
         f'font-size="13">', f'<rect width="{W}" height="{H}" fill="white"/>',
         f'<text x="{(L + W - R) / 2}" y="22" text-anchor="middle" font-size="15">{title}</text>']
    for t in nice_ticks(x0, x1):
        o.append(f'<line x1="{px(t):.1f}" y1="{T}" x2="{px(t):.1f}" y2="{H - B}" stroke="#ddd"/>')
        label = f"{t:g}" if t < 1e6 else f"{t:.3g}"
        o.append(f'<text x="{px(t):.1f}" y="{H - B + 18}" text-anchor="middle">{label}</text>')
    yticks = [10.0 ** e for e in range(int(y0), int(y1) + 1)] if logy else nice_ticks(y0, y1)
    for t in yticks:
        o.append(f'<line x1="{L}" y1="{py(t):.1f}" x2="{W - R}" y2="{py(t):.1f}" stroke="#ddd"/>')
        label = f"1e{int(round(math.log10(t)))}" if logy else f"{t:g}"
        o.append(f'<text x="{L - 8}" y="{py(t) + 4:.1f}" text-anchor="end">{label}</text>')
    o.append(f'<rect x="{L}" y="{T}" width="{W - L - R}" height="{H - T - B}" fill="none" stroke="black"/>')
    o.append(f'<text x="{(L + W - R) / 2}" y="{H - 14}" text-anchor="middle">{xlabel}</text>')
    o.append(f'<text transform="translate(20 {(T + H - B) / 2}) rotate(-90)" text-anchor="middle">{ylabel}</text>')
    for i, (name, pts) in enumerate(series):
        col = COLORS[i % len(COLORS)]
        path = " ".join(f"{px(x):.1f},{py(y):.1f}" for x, y in pts if (y > 0 or not logy))
        o.append(f'<polyline fill="none" stroke="{col}" stroke-width="1.6" points="{path}"/>')
        ly = T + 18 + 20 * i
        o.append(f'<line x1="{W - R + 12}" y1="{ly - 4}" x2="{W - R + 40}" y2="{ly - 4}" stroke="{col}" '
                 f'stroke-width="2.5"/>')
        label = name.replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;")
        o.append(f'<text x="{W - R + 46}" y="{ly}">{label}</text>')
    o.append("</svg>")
    with open(out, "w") as f:
        f.write("\n".join(o) + "\n")


def main():
    kind, path, out = sys.argv[1:4]
    header, data = read_csv(path)
    names = header[1:]
    if kind in ("gemv_flops", "dot_flops"):
        series = []
        for j, name in enumerate(names, start=1):
            if name.startswith("Error"):
                continue
            pts = []
            for r in data:
                n, ms = r[0], r[j]
                flops = 2.0 * n * n if kind == "gemv_flops" else 2.0 * n
                if ms > 0:
                    pts.append((n, flops / (ms * 1e-3) / 1e9))
            series.append((name, pts))
        svg_plot(series, "Number of rows" if kind == "gemv_flops" else "Vector size", "GFLOP/s",
                 f"{'GEMV' if kind == 'gemv_flops' else 'DOT'} on B200 (sm_100a), uniform(-1,1) data", out)
    elif kind in ("gemv_error", "dot_error", "trsv_error"):
        series = []
        for j, name in enumerate(names, start=1):
            pts = [(r[0], abs(r[j])) for r in data if abs(r[j]) > 0]
            if pts:
                series.append((name.replace("Error ", ""), pts))
        what = {"gemv_error": "GEMV relative error vs. the fp64 kernel",
                "dot_error": "DOT median relative error (10 random vector pairs) vs. the fp64 kernel",
                "trsv_error": "TRSV relative error vs. the fp64 kernel"}[kind]
        svg_plot(series, "Number of rows" if kind != "dot_error" else "Vector size", "relative error",
                 what + ", B200", out, logy=True)
    elif kind == "trsv_time":
        series = [(name, [(r[0], r[j]) for r in data]) for j, name in enumerate(names, start=1)]
        svg_plot(series, "Number of rows", "time [ms]", "TRSV on B200 (sm_100a)", out)
    else:
        raise SystemExit("unknown plot kind " + kind)


if __name__ == "__main__":
    main()
