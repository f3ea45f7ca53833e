class LinearSystem:
    pass


class StateFitting:
    pass
