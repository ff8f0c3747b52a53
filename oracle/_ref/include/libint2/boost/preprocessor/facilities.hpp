# /* **************************************************************************
#  *                                                                          *
#  *     (C) Copyright Paul Mensonides 2002-2011.                             *
#  *     (C) Copyright Edward Diener 2011.                                    *
#  *     Distributed under the Boost Software License, Version 1.0. (See      *
#  *     accompanying file LICENSE_1_0.txt or copy at                         *
#  *     http://www.boost.org/LICENSE_1_0.txt)                                *
#  *                                                                          *
#  ************************************************************************** */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_FACILITIES_HPP
# define BOOST_PREPROCESSOR_FACILITIES_HPP
#
# include <libint2/boost/preprocessor/facilities/apply.hpp>
# include <libint2/boost/preprocessor/facilities/empty.hpp>
# include <libint2/boost/preprocessor/facilities/expand.hpp>
# include <libint2/boost/preprocessor/facilities/identity.hpp>
# include <libint2/boost/preprocessor/facilities/intercept.hpp>
# include <libint2/boost/preprocessor/facilities/overload.hpp>
#
# endif
