class RayMeshIntersector:  # import stub; Embree is not available offline
    def __init__(self, mesh):
        self.mesh = mesh

    def intersects_id(self, *a, **k):
        raise NotImplementedError("Embree stand-in")
