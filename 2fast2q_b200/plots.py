"""
The four summary plots of the reference's `run_stats` (fast2q.py:1414-1527), drawn with Pillow instead of matplotlib:

    <fn>_reads_plot.png                        total / aligned / not-aligned reads per sample, horizontal bars
    <fn>_reads_plot_percentage.png             aligned | passed-but-not-aligned | quality-failed, stacked to 100 %
    <fn>_distribution_plot.png                 reads-per-feature distribution per sample (violin, quartile bar, median)
    <fn>_distribution_normalized_RPM_plot.png  the same on reads-per-million

They complete the output folder of a run (compiled.csv, compiled_stats.csv and these four: the six files
tests/test_cli.py:17-25 of the reference counts).  Same data, same colours, same file names and figure proportions
(12 in wide, rows / 4 or rows / 2 in high, 300 dpi recorded in the file); the pixels are this module's own — they are
a picture of the two csv files, not part of the count path, and nothing is compared against matplotlib's rendering.
Pillow missing -> no plots, one warning (the csv files are the result).
"""
import math

import numpy as np

DPI = 150                      # pixels per "inch" of the reference's figure sizes (its 300 would be 4x the pixels to compress)
# palette images (one byte per pixel: a third of the bytes to compress); text is drawn without anti-aliasing
PALETTE = ["#FFFFFF", "#000000", "#FFD25A", "#FFAA5A", "#F56416", "#6290C3", "#F1FFE7", "#FB5012", "#D43F3A"]
INK = {c: k for k, c in enumerate(PALETTE)}
INK.update(white=0, black=1)


def _font(px):
    from PIL import ImageFont
    try:
        return ImageFont.load_default(size=px)
    except TypeError:          # Pillow < 10.1: the fixed bitmap font
        return ImageFont.load_default()


def _nice_ticks(hi, n=6):
    """round tick positions 0 .. hi"""
    if hi <= 0:
        return [0.0]
    raw = hi / n
    mag = 10 ** math.floor(math.log10(raw))
    step = min((m for m in (1, 2, 2.5, 5, 10) if m * mag >= raw), default=10) * mag
    return [k * step for k in range(int(hi / step) + 1)]


def _fmt(v):
    if v >= 1e6:
        return f"{v / 1e6:g}M"
    if v >= 1e4:
        return f"{v / 1e3:g}k"
    return f"{v:g}"


class _Canvas:
    """a figure with one horizontal-value axis: rows on the y axis (row 0 at the bottom, like barh), values on x from xmin"""

    def __init__(self, rows, height_in, labels, xlabel, xmax, xmin=1.0, title=None, legend=None):
        from PIL import Image, ImageDraw
        self.W, self.H = 12 * DPI, max(1, int(height_in)) * DPI + 2 * DPI       # (room for the legend, title and x label)
        self.im = Image.new("P", (self.W, self.H), 0)
        self.im.putpalette([int(c[k:k + 2], 16) for c in PALETTE for k in (1, 3, 5)])
        self.d = ImageDraw.Draw(self.im)
        self.d.fontmode = "1"
        self.tick_font, self.label_font, self.legend_font = _font(DPI // 7), _font(DPI // 5), _font(DPI // 9)
        wlab = max([self.d.textlength(str(s), font=self.tick_font) for s in labels] + [10])
        self.x0, self.x1 = int(wlab) + DPI // 4, self.W - DPI // 3
        self.y0, self.y1 = self.H - DPI, DPI // 2 + (DPI // 3 if title else 0) + (DPI // 3 if legend else 0)
        self.rows, self.xmin, self.xmax = rows, xmin, max(xmax, xmin + 1.0)
        self.row_h = (self.y0 - self.y1) / max(rows, 1)
        # axes (left and bottom spines only), ticks, labels
        self.d.line([(self.x0, self.y0), (self.x1, self.y0)], fill=1, width=2)
        self.d.line([(self.x0, self.y0), (self.x0, self.y1)], fill=1, width=2)
        for t in _nice_ticks(self.xmax):
            if t < self.xmin and t != 0:
                continue
            x = self.x(max(t, self.xmin))
            self.d.line([(x, self.y0), (x, self.y0 + DPI // 20)], fill=1, width=2)
            s = _fmt(t)
            self.d.text((x - self.d.textlength(s, font=self.tick_font) / 2, self.y0 + DPI // 14), s, font=self.tick_font, fill=1)
        for r, s in enumerate(labels):
            y = self.y(r)
            self.d.line([(self.x0 - DPI // 20, y), (self.x0, y)], fill=1, width=2)
            self.d.text((self.x0 - DPI // 12 - self.d.textlength(str(s), font=self.tick_font), y - DPI // 12), str(s), font=self.tick_font, fill=1)
        self.d.text(((self.x0 + self.x1) / 2 - self.d.textlength(xlabel, font=self.label_font) / 2, self.y0 + DPI // 3), xlabel,
                    font=self.label_font, fill=1)
        ytop = DPI // 8
        if title:
            self.d.text(((self.x0 + self.x1) / 2 - self.d.textlength(title, font=self.label_font) / 2, ytop), title, font=self.label_font, fill=1)
            ytop += DPI // 3
        if legend:
            x = self.x0
            for text, colour in legend:
                self.d.rectangle([x, ytop, x + DPI // 5, ytop + DPI // 8], fill=INK[colour], outline=1)
                self.d.text((x + DPI // 4, ytop - DPI // 60), text, font=self.legend_font, fill=1)
                x += DPI // 4 + int(self.d.textlength(text, font=self.legend_font)) + DPI // 4

    def x(self, v):
        return self.x0 + (self.x1 - self.x0) * (v - self.xmin) / (self.xmax - self.xmin)

    def y(self, row):
        """centre of a row"""
        return self.y0 - (row + 0.5) * self.row_h

    def bar(self, row, left, right, colour, width=0.75):
        if right <= max(left, self.xmin):
            return
        h = self.row_h * width / 2
        self.d.rectangle([self.x(max(left, self.xmin)), self.y(row) - h, self.x(min(right, self.xmax)), self.y(row) + h], fill=INK[colour], outline=1)

    def save(self, path):
        """encoded on a worker thread (the encoder releases the GIL); write_all joins them"""
        import threading
        t = threading.Thread(target=self.im.save, args=(path,), kwargs=dict(dpi=(300, 300), compress_level=1))
        t.start()
        _pending.append(t)


_pending = []


def _stat_rows(table):
    """the per-sample rows of <fn>_stats.csv (behind the '#Sample name' header row)"""
    for k, row in enumerate(table):
        if row and row[0] == "#Sample name":
            return table[k + 1:], len(table)
    return [], len(table)


def reads_plots(table, directory_prefix):
    """the two bar plots from the rows of <fn>_stats.csv"""
    rows, n_all = _stat_rows(table)
    if not rows:
        return
    names = [r[0] for r in rows]
    total = [int(r[3]) for r in rows]
    aligned = [int(r[4]) for r in rows]
    not_aligned = [int(r[7]) for r in rows]
    q_failed = [int(r[8]) for r in rows]
    c = _Canvas(len(rows), n_all / 4, names, "Number of reads", max(total + [1]) * 1.05,
                legend=[("Total reads in sample", "#FFD25A"), ("Aligned reads", "#FFAA5A"),
                        ("Reads that passed quality filtering but failed to align", "#F56416")])
    for i in range(len(rows)):                                # drawn over each other, as the reference does
        c.bar(i, 0, total[i], "#FFD25A"); c.bar(i, 0, aligned[i], "#FFAA5A"); c.bar(i, 0, not_aligned[i], "#F56416")
    c.save(directory_prefix + "_reads_plot.png")
    c = _Canvas(len(rows), n_all / 4, names, "% of reads per sample", 105.0,
                legend=[("Aligned reads", "#6290C3"), ("Reads that passed quality filtering but failed to align", "#F1FFE7"),
                        ("Reads that did not pass quality filtering", "#FB5012")])
    for i in range(len(rows)):
        t = max(total[i], 1)
        a, n, q = aligned[i] / t * 100, not_aligned[i] / t * 100, q_failed[i] / t * 100
        c.bar(i, 0, a, "#6290C3"); c.bar(i, a, a + n, "#F1FFE7"); c.bar(i, a + n, a + n + q, "#FB5012")
    c.save(directory_prefix + "_reads_plot_percentage.png")


def _kde(values, points=200):
    """gaussian kernel density on `points` positions between min and max (Scott's rule), from a histogram of the values
    (30 000 features x 200 positions per sample would be the slow way)"""
    v = np.asarray(values, dtype=np.float64)
    lo, hi = float(v.min()), float(v.max())
    xs = np.linspace(lo, hi, points)
    sd = float(v.std(ddof=1)) if v.size > 1 else 0.0
    if sd == 0.0 or hi == lo:
        return xs, np.ones(points)
    bw = sd * v.size ** (-0.2)
    bins = min(4096, max(points, v.size))
    hist, edges = np.histogram(v, bins=bins, range=(lo, hi))
    centres = (edges[:-1] + edges[1:]) / 2
    nz = hist > 0
    dens = (hist[nz][None, :] * np.exp(-0.5 * ((xs[:, None] - centres[nz][None, :]) / bw) ** 2)).sum(axis=1)
    return xs, dens


def distribution_plots(head, compiled, n_stat_rows, directory_prefix):
    """the two violin plots from the columns of <fn>.csv"""
    samples = head[1:]
    if not samples or not compiled:
        return
    data = np.asarray([compiled[f] for f in compiled], dtype=np.float64).T            # [sample, feature]
    sets = [("Reads per feature distribution", "_distribution_plot.png", list(data), samples)]
    keep = [k for k in range(len(samples)) if data[k].sum() > 0]
    if keep:
        sets.append(("Reads per feature (RPM normalized) distribution", "_distribution_normalized_RPM_plot.png",
                     [data[k] / data[k].sum() * 1e6 for k in keep], samples))         # (labels as in the reference: all names)
    for title, suffix, rows, labels in sets:
        xmax = max(float(r.max()) for r in rows) * 1.05
        c = _Canvas(len(labels), n_stat_rows / 2, labels, "Reads per feature", xmax, title=title)
        for i, r in enumerate(rows):
            xs, dens = _kde(r)
            half = 0.5 * dens / dens.max() * c.row_h
            top = [(c.x(max(x, c.xmin)), c.y(i) - h) for x, h in zip(xs, half)]
            bot = [(c.x(max(x, c.xmin)), c.y(i) + h) for x, h in zip(xs[::-1], half[::-1])]
            c.d.polygon(top + bot, fill=INK["#D43F3A"], outline=1)
            q1, med, q3 = np.percentile(r, [25, 50, 75])
            c.d.line([(c.x(max(q1, c.xmin)), c.y(i)), (c.x(max(q3, c.xmin)), c.y(i))], fill=1, width=max(2, DPI // 18))
            xm, rad = c.x(max(med, c.xmin)), DPI // 24
            c.d.ellipse([xm - rad, c.y(i) - rad, xm + rad, c.y(i) + rad], fill=0, outline=1)
        c.save(directory_prefix + suffix)


def write_all(table, head, compiled, directory_prefix):
    """all four; returns the list of files written ([] when Pillow is missing)"""
    try:
        import PIL  # noqa: F401
    except ImportError:
        return None
    reads_plots(table, directory_prefix)
    distribution_plots(head, compiled, len(table), directory_prefix)
    while _pending:
        _pending.pop().join()
    return [directory_prefix + s for s in ("_reads_plot.png", "_reads_plot_percentage.png", "_distribution_plot.png",
                                           "_distribution_normalized_RPM_plot.png")]
