"""Small host helpers that sit on the annotation path (reference cta/utils.py:16-146).  The crop /
soft-mask functions of the reference (utils.py:226-270) live in csrc/stage3_patches.cu."""
import colorsys

VOTE_ORDER = ["CD4 T cell", "CD8 T cell", "Dendritic cell", "B cell", "M1 macrophage cell", "M2 macrophage cell",
              "Regulatory T cell", "Granulocyte cell", "Plasma cell", "Natural killer cell", "Mast cell",
              "Stroma cell", "Smooth muscle", "Endothelial cell", "Epithelial cell", "Proliferating/tumor cell",
              "Nerve cell"]


def get_void_vote():
    """Key order = tie-break order of the multi-model merge (utils.py:143-146)."""
    return {k: 0 for k in VOTE_ORDER}


def get_colors(n):
    """n visually distinct RGB triples (presentation only)."""
    return [tuple(int(255 * c) for c in colorsys.hsv_to_rgb(i / max(n, 1), 0.65, 0.95)) for i in range(n)]


def rgb_to_hex(rgb):
    return "#{:02x}{:02x}{:02x}".format(*(int(v) for v in rgb))


def number_to_rgb(x):
    """confidence in [0, 1] -> blue..red ramp (presentation only)."""
    x = min(max(float(x), 0.0), 1.0)
    return [int(255 * x), 64, int(255 * (1 - x))]
