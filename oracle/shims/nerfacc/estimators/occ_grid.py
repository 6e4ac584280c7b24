class OccGridEstimator:  # import stub
    pass
