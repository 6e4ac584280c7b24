class PropNetEstimator:  # import stub
    pass
