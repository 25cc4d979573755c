# Unit constants of the reference (src/jaxoplanet/constants.py:1-9): R_sun, M_sun, day.
# c.G.to("Rsun3/Msun.day2").value
G = 2942.2062175044193
# (1 * u.au).to(u.R_sun).value
au = 215.03215567054764
M_earth = 3.0034893488507934e-06
# (1 * u.Mjup).to("Msun").value
M_jup = 0.0009545942339693249
# (1 * u.m / u.s).to(u.Rsun / u.day).value
m_per_s = 0.00012419146183699872
