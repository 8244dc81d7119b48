def marching_cubes(*a, **k):
    raise NotImplementedError('skimage.measure.marching_cubes is not available in the test sandbox')
