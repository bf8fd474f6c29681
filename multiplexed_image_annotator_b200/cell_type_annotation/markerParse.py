"""Marker list -> per-panel channel indices.  Host-side, tiny; decides which kernels and models run.

Mirrors reference cta/markerParse.py:4-117 (`MarkerParser`): five fixed antibody panels, alias
replacement on the fixed-width numpy string array np.loadtxt returns (so a long alias is truncated,
quirk Q10), per-panel missing-marker budgets when `strict` is False, `None` for a panel that is not
applied, and the five boolean flags the Annotator branches on.
"""
import numpy as np

PANEL_TABLE = (
    # key, flag attribute, missing budget, markers (channel order the classifier was trained on)
    ("immune_base", "immune_base", 1, ("CD45", "CD20", "CD4", "CD8", "DAPI", "CD11c", "CD3")),
    ("immune_extended", "immune_extended", 2,
     ("DAPI", "CD3", "CD4", "CD8", "CD11c", "CD20", "CD45", "CD68", "CD163", "CD56")),
    ("immune_full", "immune_full", 3,
     ("DAPI", "CD3", "CD4", "CD8", "CD11c", "CD15", "CD20", "CD45", "CD56", "CD68", "CD138", "CD163", "FoxP3",
      "Granzyme B", "Trypase")),          # sic: the reference spells it "Trypase" (Q12)
    ("structure", "struct", 1, ("DAPI", "aSMA", "CD31", "PanCK", "Vimentin", "Ki67", "CD45")),
    ("nerve_cell", "nerve", 0, ("DAPI", "CD45", "GFAP")),
)
ALIASES = {"DNA": "DAPI", "DPAI-02": "DAPI", "CD16": "CD15", "CD38": "CD138", "CD79": "CD20", "CHGA": "GFAP",
           "SMActin": "aSMA", "CD3e": "CD3", "CK": "PanCK", "CytoKeratin": "PanCK", "Cytokeratin": "PanCK",
           "Cytokeratin-19": "PanCK", "panCK": "PanCK"}
ALTERNATIVES = {"CD20": "CD20 or CD79a", "GFAP": "GFAP or Chromogranin A", "CD138": "CD138 or CD38"}


class MarkerParser:
    def __init__(self, strict=True, logger=None):
        self.panels = {key: list(markers) for key, _, _, markers in PANEL_TABLE}
        self.indices = {}
        for _, flag, _, _ in PANEL_TABLE:
            setattr(self, flag, False)
        self.strict = strict
        self.markers = []
        self.logger = logger

    def _say(self, text, echo=False):
        if echo:
            print(text, end="")
        if self.logger:
            self.logger.log(text.rstrip(", "))

    def _matching(self, marker_list, panel, panel_name):
        budget = {key: b for key, _, b, _ in PANEL_TABLE}[panel_name]
        matched, missing = [], []
        for marker in panel:
            if marker in marker_list:
                matched.append(marker_list.index(marker))
                continue
            shown = ALTERNATIVES.get(marker, marker)
            if self.strict or len(panel) <= 3:
                self._say(f"Marker {shown} is not found in the list, ", echo=True)
                return None
            missing.append(shown)
            matched.append(-1)
            if len(missing) > budget:
                self._say(f"Markers {', '.join(missing)} are not found in the list, ", echo=True)
                return None
        return matched

    def parse(self, marker_file):
        names = np.loadtxt(marker_file, delimiter=",", dtype=str)       # fixed-width <U array
        if names.ndim == 0:
            raise TypeError("iteration over a 0-d array")               # same failure as the reference on one line
        self.markers.extend(names)
        if self.logger:
            self.logger.log("The panel contains the following markers: " + ", ".join(names) + ".")
        for i in range(len(names)):
            alias = ALIASES.get(str(names[i]))
            if alias is not None and alias not in names:
                old = names[i]
                names[i] = alias                                        # may truncate (Q10)
                if self.logger:
                    self.logger.log(f"Replaced the marker name {old} with {names[i]} to match our panel.")
        if self.logger:
            self.logger.log("")
        marker_list = list(names)
        self.n_markers = len(marker_list)
        for key, flag, _, markers in PANEL_TABLE:
            matched = self._matching(marker_list, list(markers), key)
            applied = bool(matched)
            self.indices[key] = matched if applied else None
            setattr(self, flag, applied)
            state = "applied" if applied else "not applied"
            print(f"{key} panel is {state}.")
            if self.logger:
                self.logger.log(f"{key} panel is {state}.")
                self.logger.log("\n")
