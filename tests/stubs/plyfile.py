"""Stand-in for the plyfile calls of utils/mesh_util.py:60-74 (test infrastructure): ASCII PLY writer."""
import numpy as np


class PlyElement:
    def __init__(self, data, name):
        self.data, self.name = data, name

    @staticmethod
    def describe(data, name):
        return PlyElement(np.asarray(data), name)


class PlyData:
    def __init__(self, elements):
        self.elements = elements

    def write(self, path):
        with open(path, 'w') as f:
            f.write('ply\nformat ascii 1.0\n')
            for e in self.elements:
                f.write(f'element {e.name} {len(e.data)}\n')
                for n in (e.data.dtype.names or ()):
                    f.write(f'property float {n}\n')
            f.write('end_header\n')
            for e in self.elements:
                for row in e.data:
                    f.write(' '.join(str(x) for x in np.asarray(row.tolist()).reshape(-1)) + '\n')
