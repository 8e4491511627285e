class TimeLimit:
    pass
