class Figure: pass
class Heatmap: pass
