"""`skspatial.objects.Line` stand-in: `project_point` is the orthogonal projection onto the line,
point + direction * ((p - point) . direction) / (direction . direction) -- scikit-spatial's documented formula."""
import numpy as np


class Line:
    def __init__(self, point, direction):
        self.point = np.asarray(point, dtype=np.float64)
        self.direction = np.asarray(direction, dtype=np.float64)

    def project_point(self, p):
        p = np.asarray(p, dtype=np.float64)
        return self.point + self.direction * (np.dot(p - self.point, self.direction) / np.dot(self.direction, self.direction))
